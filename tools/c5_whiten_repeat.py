"""Config-5 whitening, repeated: wall time of every call with the library's per-phase log (B200LM_VERBOSE) -- looks
for calls that leave the fast one-sided path."""
import os
import sys
import time

os.environ["B200LM_VERBOSE"] = "1"
import numpy as np
import torch

sys.path.insert(0, ".")
from lsqfit_b200 import configs
from lsqfit_b200.whiten import whiten_blocks

cfg = configs.c5()
flat = cfg["ycov"].reshape(-1)
for i in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    W, Cc, nout, nmod, logdet = whiten_blocks([cfg["ny"]], flat, cfg["svdcut"], None, 0, as_torch=True)
    torch.cuda.synchronize()
    print("CALL %d: %.3f s nmod %d logdet %.6f" % (i, time.perf_counter() - t0, int(nmod[0]), float(logdet[0])), file=sys.stderr, flush=True)
    del W, Cc
