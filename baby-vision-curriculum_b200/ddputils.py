"""Host-side mirror of pretraining/generative/ddputils.py:53-68: the scalar-loss all-reduce the reference loop applies
to `outputs.loss` (pretrain_videomae.py:303).  Forward: x / world_size then all_reduce(SUM) (= mean over ranks, a
logging value); backward: identity (DDP averages the gradients separately)."""
import torch
import torch.distributed as dist


class AllReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            x = x.contiguous() / dist.get_world_size()
            dist.all_reduce(x)
        return x

    @staticmethod
    def backward(ctx, grads):
        return grads
