"""Data-parallel gradient exchange: one all-reduce over the flat gradient arena.

Replaces the in-graph clone sum of /root/reference/slim/deployment/model_deploy.py:414-444
(`_sum_clones_gradients`: tf.add_n on the CPU) and its loss scaling (:221-225 clone loss / num_clones,
:296 regularisation losses added once).  With torch.distributed the backend is NCCL over
NVLink-5 / NVSwitch on the GPUs and gloo in the CPU tests."""
import torch
import torch.distributed as dist


def data_parallel_scale(world_size):
    """Factor applied to the summed task-loss gradients (the L2 gradient is added after scaling)."""
    return 1.0 / float(world_size)


def allreduce_gradients(flat_grads, world_size, group=None, async_op=False):
    """Sum one gradient bucket (a contiguous view of the gradient arena) over the replicas, in place.
    async_op=True returns the work handle (None with a single replica): the caller overlaps the exchange with compute
    and calls `.wait()` on the stream that consumes the sum."""
    if world_size > 1:
        work = dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return work if async_op else flat_grads
    return None if async_op else flat_grads


def average_losses(task_losses, world_size, group=None):
    """The logged task losses of a data-parallel step: sum over the replicas / num_replicas, in place (what the
    reference's TotalLoss is made of: every clone's loss is divided by num_clones and the clones are summed,
    model_deploy.py:221-225).  The regularisation term is identical on every replica and is not exchanged."""
    if world_size > 1:
        dist.all_reduce(task_losses, op=dist.ReduceOp.SUM, group=group)
        task_losses.mul_(1.0 / float(world_size))
    return task_losses
