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


class AllGather(torch.autograd.Function):
    """pretraining/predictive/distributed.py:49-76 (defined by the reference, unused by its loops; BASELINE.json
    config 3 -- NT-Xent over the embeddings of ALL ranks -- is built from it): forward concatenates every rank's x
    along dim 0 in rank order; backward sums the incoming gradient over the ranks and returns this rank's rows.

        feats = bvc_b200.AllGather.apply(local_feats)                 # [world * 2B, D]
        loss = bvc_b200.info_nce_loss(temperature, masks_for(world * 2B), feats)

    One collective writes straight into the concatenated tensor; on NCCL the backward is a reduce-scatter (each rank
    receives only its own rows) instead of the reference's all-reduce of the whole gradient."""

    @staticmethod
    def forward(ctx, x):
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            ctx.world = 1
            return x
        world = dist.get_world_size()
        ctx.world, ctx.rows = world, x.shape[0]
        x = x.contiguous()
        out = x.new_empty((world * x.shape[0],) + tuple(x.shape[1:]))
        dist.all_gather(list(out.chunk(world, dim=0)), x)
        return out

    @staticmethod
    def backward(ctx, grads):
        if ctx.world == 1:
            return grads
        grads = grads.contiguous()
        if dist.get_backend() == "nccl":
            mine = grads.new_empty((ctx.rows,) + tuple(grads.shape[1:]))
            dist.reduce_scatter_tensor(mine, grads)
            return mine
        dist.all_reduce(grads)
        r = dist.get_rank()
        return grads[r * ctx.rows:(r + 1) * ctx.rows]
