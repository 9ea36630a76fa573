"""Walker sharding over the GPUs of one box.

Walkers (parameter rows) are independent -- the reference evaluates them in separate processes
(examples/Sun/01_emcee.py:119-122) -- so a batch is block-partitioned over the ranks of a
``torch.distributed`` group, each rank evaluates its rows on its own GPU, and ONE collective
gathers the per-walker log-likelihoods (NCCL on GPUs; the same code runs on gloo / CPU tensors
in the tests with an injected evaluator).  No matrix is ever split across GPUs and sharding does
not change any value: every rank produces bit-identical numbers to a single-GPU run.
"""
import numpy as np


def shard_bounds(W, world_size, rank):
    """[lo, hi) rows of rank `rank`: contiguous blocks, the first W % world_size ranks get one more."""
    base, rem = divmod(int(W), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(W, world_size):
    return [shard_bounds(W, world_size, r)[1] - shard_bounds(W, world_size, r)[0] for r in range(world_size)]


def gather_loglikes(local, W, group=None):
    """All-gather the per-rank result slices (1-D float64 torch tensors, sizes = shard_sizes) into the
    full [W] vector on every rank.  One collective; uneven shards are padded to the largest."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(W, world)
    mx = max(sizes)
    if local.numel() != sizes[dist.get_rank(group)]:
        raise ValueError("local slice has %d rows, expected %d" % (local.numel(), sizes[dist.get_rank(group)]))
    if min(sizes) == mx:
        out = torch.empty(W, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.full((mx,), float("nan"), dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    buf = torch.empty(mx * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * mx:r * mx + sizes[r]] for r in range(world)])


def init_comm(net, group=None):
    """Bind an NCCL communicator over the ranks of `group` to the network's context (agpn_comm_init): rank 0 draws
    the ncclUniqueId inside the library, torch.distributed only carries its 128 bytes to the other ranks.  Idempotent
    per (network, group)."""
    import torch.distributed as dist
    from . import _lib
    key = ("comm", id(group), dist.get_world_size(group), dist.get_rank(group))
    if getattr(net, "_comm_key", None) == key and net._handle is not None:
        return
    lib = _lib.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    box = [None]
    if rank == 0:
        buf = np.zeros(128, dtype=np.uint8)
        _lib.check(lib.agpn_comm_unique_id(_lib.ptr(buf)))
        box[0] = buf.tobytes()
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast_object_list(box, src=src, group=group)
    idb = np.frombuffer(box[0], dtype=np.uint8).copy()
    _lib.check(lib.agpn_comm_init(net._ctx(), _lib.ptr(idb), rank, world))
    net._comm_key = key


def log_likelihood_batch_sharded(net, theta, group=None, evaluator=None):
    """Log-likelihoods of all W rows of `theta` (same full array on every rank), computed
    W / world_size rows per rank and gathered.

    Default (GPU): one call into the C ABI (agpn_loglike_batch_sharded): the rank's rows are evaluated with the
    results written into the gather buffer and one ncclAllGather on the current stream returns the [W] vector --
    torch only hands over buffers.  `evaluator(theta_slice) -> 1-D float64 torch tensor` replaces the GPU evaluation
    and gathers through torch.distributed instead (CPU / gloo tests of the partitioning logic)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    W = theta.shape[0]
    if evaluator is not None:
        lo, hi = shard_bounds(W, world, rank)
        return gather_loglikes(evaluator(theta[lo:hi]), W, group)
    from . import _lib
    if net._structure is None:
        raise RuntimeError("call set_structure(nodes, weights, means) first")
    init_comm(net, group)
    lib = _lib.load()
    dev = torch.device("cuda", net._device)
    if not torch.is_tensor(theta):
        theta = torch.from_numpy(np.ascontiguousarray(theta, dtype=np.float64)).to(dev)
    if theta.dtype != torch.float64 or theta.dim() != 2 or theta.shape[1] != net._structure.D or theta.device != dev:
        raise ValueError("theta must be float64 [W, %d] on cuda:%d" % (net._structure.D, net._device))
    theta = theta.contiguous()
    out = torch.empty(W, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.agpn_loglike_batch_sharded(net._ctx(), W, theta.data_ptr(), 1, out.data_ptr(), 1, None, stream))
    return out


def prediction_sharded(net, nodes, weights, means, jitters, time, dataset=1, group=None):
    """network.prediction with the prediction grid split over the ranks: every rank
    factorises the same training block (cheap next to the N^2 M solve), predicts its contiguous slice of
    `time` and one all-gather returns the full (mean, std, time) on every rank.  mean.Linear / Keplerian
    see the statistics of the WHOLE grid.  Restrictions: see agpn_predict_slice."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    time = np.ascontiguousarray(time, dtype=np.float64)
    M = time.size
    lo, hi = shard_bounds(M, world, rank)
    mu, sd = net.prediction_slice(nodes, weights, means, jitters, time, lo, hi, dataset)
    dev = torch.device("cuda", net._device)
    both = torch.from_numpy(np.concatenate([mu, sd])).to(dev) if hi > lo else torch.empty(0, dtype=torch.float64, device=dev)
    sizes = shard_sizes(M, world)
    mx = max(sizes)
    pad = torch.full((2 * mx,), float("nan"), dtype=torch.float64, device=dev)
    n = hi - lo
    pad[:n] = both[:n]
    pad[mx:mx + n] = both[n:]
    buf = torch.empty(2 * mx * world, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(buf, pad, group=group)
    buf = buf.cpu().numpy().reshape(world, 2, mx)
    y_mean = np.concatenate([buf[r, 0, :sizes[r]] for r in range(world)])
    y_std = np.concatenate([buf[r, 1, :sizes[r]] for r in range(world)])
    return y_mean, y_std, time
