"""2+ GPU probe: cost of the per-step NCCL all_gather of the log-likelihoods (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from artgpn_b200 import synth, node, weight, mean, sharding
from artgpn_b200.arttwo import network
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t, y, sig = synth.make_data(2048, 3, seed=0)
theta = synth.make_walkers(W * world, p=3, q=2, seed=1)
net = network(2, t, y[0], sig[0], y[1], sig[1], y[2], sig[2], device=lr)
net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], 3, 2)[:3])
lo, hi = sharding.shard_bounds(W * world, world, rank)
th = torch.from_numpy(np.ascontiguousarray(theta[lo:hi])).to(dev)
def timed(fn, n=5):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
loc = net.log_likelihood_batch(th)
a = timed(lambda: net.log_likelihood_batch(th))
b = timed(lambda: sharding.gather_loglikes(net.log_likelihood_batch(th), W * world))
c = timed(lambda: sharding.gather_loglikes(loc, W * world), 50)
print("rank %d W/gpu=%d compute %.3f ms  compute+gather %.3f ms  gather only %.3f ms" % (rank, W, a, b, c), file=sys.stderr)
dist.barrier(); dist.destroy_process_group()
