"""TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is the CPU checker for the B200 hot path: import shims that let the
*unmodified* reference modules load, a loader for those modules, seeded synthetic inputs/weights,
and a plain torch-fp32 / numpy restatement of the reference algorithm that can travel to the GPU
box (where ``/root/reference`` does not exist).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product (``unet-zoo_b200/``) never does.
"""
