"""B200 (sm_100a) execution layer behind the UNet-Zoo drop-in modules.

``_lib``  : ctypes binding of libunetzoo_b200.so (prototypes parsed from include/unetzoo_b200.h)
``ops``   : torch.autograd.Functions that own saved tensors / workspaces and call the C ABI
``dp``    : one-process-per-GPU data parallelism (NCCL gradient all-reduce, sample-sharded evaluation)

There is no CPU or PyTorch fallback: every op raises if the library is missing or the tensor is not on a CUDA device.
"""
