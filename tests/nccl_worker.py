"""Worker of tests/test_gpu_round2.py::test_nccl_ranks_match_one_device: launched by torchrun, one rank per GPU.
Every rank evaluates its run of pairwise-tree chunks, ONE NCCL all_gather exchanges the ~1 KB chunk records, every
rank folds them in tree order and must hold the bit pattern one device produces alone."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_trade_sizing_b200 import Engine, make_params  # noqa: E402
from mc_trade_sizing_b200 import distributed as mdist  # noqa: E402
from mc_trade_sizing_b200 import trading_models as tm  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
eng = Engine(local)
ok = True
for args, kw in (((0.82, 186.0, 219.0, 1000, 1 << 15, 20), {}),                                    # 655,360 paths, fp64 evaluator
                 ((0.82, 186.0, 219.0, 1000, 40000, 7), dict(use_markov=True, p_win_after_win=0.75, p_win_after_loss=0.6)),   # int32 evaluator
                 ((0.82, 186.0, 219.0, 64, 130, 7), {})):                                            # 910 paths: fewer chunks than some world sizes
    prm = make_params(*args, **kw)
    rows = mdist.run_all_strategies_sharded(eng, prm, 77, rank, world, device=device)               # chunk size chosen by resolve_chunk
    one = tm.summary_to_rows(eng.simulate(prm, seed=77)["summary"])
    same = [r[0] for r in rows] == [r[0] for r in one] and \
        (np.array([r[1:] for r in rows], dtype=np.float64).view(np.uint64) == np.array([r[1:] for r in one], dtype=np.float64).view(np.uint64)).all()
    ok = ok and bool(same)
flag = torch.tensor([1 if ok else 0], device=device)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("NCCL_SHARDED_OK" if int(flag[0]) == 1 else "NCCL_SHARDED_MISMATCH", "world", world, flush=True)
dist.barrier()
dist.destroy_process_group()
