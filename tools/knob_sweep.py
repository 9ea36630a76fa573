"""Sweep of scheduling knobs (sub-batch streams, update clusters, look-ahead) at the BASELINE config-3 shape."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from artgpn_b200 import synth, node, weight, mean
from artgpn_b200.arttwo import network
N, p, q = 2048, 3, 2
t, y, sig = synth.make_data(N, p, seed=0)
KNOBS = ("AGPN_STREAMS", "AGPN_I8_CLUSTERS", "AGPN_LOOKAHEAD", "AGPN_TRSM_SPLIT")
def run(W, env, reps=20):
    for k in KNOBS: os.environ.pop(k, None)
    os.environ.update(env)
    theta = synth.make_walkers(W, p=p, q=q, seed=1)
    net = network(q, t, y[0], sig[0], y[1], sig[1], y[2], sig[2])
    net.set_structure(*synth.objects_from_row(node, weight, mean, theta[0], p, q)[:3])
    th = torch.from_numpy(theta).cuda()
    for _ in range(4): net.log_likelihood_batch(th)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): net.log_likelihood_batch(th)
    b.record(); torch.cuda.synchronize()
    net.close()
    return a.elapsed_time(b) / reps
SWEEP = ((128, [{}, {"AGPN_STREAMS": "2"}, {"AGPN_STREAMS": "3"}, {"AGPN_STREAMS": "4"}, {"AGPN_STREAMS": "6"}, {"AGPN_STREAMS": "8"},
                {"AGPN_TRSM_SPLIT": "0"}]),
         (64, [{}, {"AGPN_STREAMS": "3"}, {"AGPN_STREAMS": "6"}, {"AGPN_STREAMS": "8"}]),
         (16, [{}, {"AGPN_STREAMS": "3"}, {"AGPN_STREAMS": "4"}, {"AGPN_STREAMS": "8"}, {"AGPN_I8_CLUSTERS": "16"}, {"AGPN_I8_CLUSTERS": "37"},
               {"AGPN_I8_CLUSTERS": "74"}, {"AGPN_LOOKAHEAD": "2"}, {"AGPN_TRSM_SPLIT": "0"}]))
if len(sys.argv) > 1 and sys.argv[1] == "split":
    SWEEP = tuple((W, [{}, {"AGPN_TRSM_SPLIT": "0"}, {"AGPN_TRSM_SPLIT": "1"}, {"AGPN_LOOKAHEAD": "2"}, {"AGPN_LOOKAHEAD": "0"}])
                  for W in (8, 16, 32, 48, 64, 96, 128))
for W, envs in SWEEP:
    for env in envs:
        print("W=%d %s: %.3f ms" % (W, env, run(W, env)), flush=True)
