"""Restatement of ``revtorch==0.2.0`` (reference requirements.txt:38), absent from the image.

Call sites in the reference: torchlayers.py:4,75,78 and models/phiseg3D.py:5,81,84.
Algorithm (RevNet additive coupling, Gomez et al. 2017; revtorch README):
    forward :  x1,x2 = chunk(x,2,dim);  y1 = x1 + F(x2);  y2 = x2 + G(y1);  y = cat(y1,y2)
    backward:  gy1 = G(y1) (re-run, grad enabled) ; x2 = y2 - gy1 ; dx1 = dy1 + d(gy1.dy2)/dy1
               fx2 = F(x2) (re-run, grad enabled) ; x1 = y1 - fx2 ; dx2 = dy2 + d(fx2.dx1)/dx2
Only block outputs are kept in forward; F and G therefore run twice per training step (BatchNorm
running statistics receive two momentum updates).  Sub-module names ``f_block`` / ``g_block`` /
``reversible_blocks`` define the reference state_dict keys.

PARITY UNPINNED: the reference holds no test for this package and its source is not in the container;
this file follows the published algorithm.
"""
import torch
import torch.nn as nn


class ReversibleBlock(nn.Module):
    def __init__(self, f_block, g_block, split_along_dim=1, fix_random_seed=False):
        super().__init__()
        self.f_block = f_block
        self.g_block = g_block
        self.split_along_dim = split_along_dim

    def forward(self, x):
        x1, x2 = torch.chunk(x, 2, dim=self.split_along_dim)
        with torch.no_grad():
            y1 = x1 + self.f_block(x2)
            y2 = x2 + self.g_block(y1)
        return torch.cat([y1, y2], dim=self.split_along_dim)

    def backward_pass(self, y, dy, retain_graph=False):
        y1, y2 = torch.chunk(y, 2, dim=self.split_along_dim)
        dy1, dy2 = torch.chunk(dy, 2, dim=self.split_along_dim)
        y1 = y1.detach().requires_grad_(True)
        with torch.enable_grad():
            gy1 = self.g_block(y1)
            gy1.backward(dy2, retain_graph=retain_graph)
        with torch.no_grad():
            x2 = y2 - gy1
            dx1 = dy1 + y1.grad
        x2 = x2.detach().requires_grad_(True)
        with torch.enable_grad():
            fx2 = self.f_block(x2)
            fx2.backward(dx1, retain_graph=retain_graph)
        with torch.no_grad():
            x1 = y1 - fx2
            dx2 = dy2 + x2.grad
            x = torch.cat([x1, x2.detach()], dim=self.split_along_dim)
            dx = torch.cat([dx1, dx2], dim=self.split_along_dim)
        return x, dx


class _ReversibleModuleFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, reversible_blocks, *params):
        with torch.no_grad():
            for block in reversible_blocks:
                x = block(x)
        ctx.y = x.detach()
        ctx.reversible_blocks = reversible_blocks
        return x

    @staticmethod
    def backward(ctx, dy):
        y = ctx.y
        del ctx.y
        for block in ctx.reversible_blocks[::-1]:
            y, dy = block.backward_pass(y, dy)
        return (dy, None) + tuple(None for _ in ctx.reversible_blocks.parameters())


class ReversibleSequence(nn.Module):
    def __init__(self, reversible_blocks, eagerly_discard_variables=True):
        super().__init__()
        self.reversible_blocks = reversible_blocks

    def forward(self, x):
        if torch.is_grad_enabled():
            # parameters are passed so autograd knows the output needs grad even if x does not
            return _ReversibleModuleFunction.apply(x, self.reversible_blocks, *self.reversible_blocks.parameters())
        for block in self.reversible_blocks:
            x = block(x)
        return x
