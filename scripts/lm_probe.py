import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import towr_b200 as tb
from towr_b200.configs import synthetic_iterates_fast
from towr_b200.solver import BatchedLevenbergMarquardt
p = tb.Problem(tb.make_formulation("anymal_trot_block").to_spec())
B = 4096
bt = p.batch(B)
lm = BatchedLevenbergMarquardt(bt)
lm.run(torch.from_numpy(synthetic_iterates_fast(p, B, seed=3)).cuda(), 2)
torch.cuda.synchronize()
