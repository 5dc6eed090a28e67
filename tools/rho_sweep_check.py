"""Replicas-only rho_test sweep (opt.py:882-910) across the GPUs of one box: rank r takes minibatches j % world == r, no
collective on the data path.  Launched with torch.distributed.run; rank 0 compares the gathered rows with its own
sequential sweep over all minibatches and prints both wall times."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optwboundeigenval_b200 import hvp_operator, zoo                    # noqa: E402
from optwboundeigenval_b200.spectral import SpectralState               # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "cifar_densenet"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
model, loss = zoo.build(kind)
model.train()
batch = zoo.CONFIGS[kind][3]
loader = [zoo.synthetic_batch(kind, batch, seed=zoo.SEED + 17 * k) for k in range(nb)]
st = SpectralState(model, loss, pow_iter_eps=0.0, max_pow_iter=20, ignore_bad_vals=False, rand_init=True)
st.rho_test(loader[:world])                                            # warm-up: plans, graphs
dist.barrier(); torch.cuda.synchronize()
t0 = time.time()
stats, avg = st.rho_test(loader)
torch.cuda.synchronize(); dist.barrier()
t_sharded = time.time() - t0
if rank == 0:
    hvp_operator.set_data_parallel(False)
    st.comp_rho(loader[0])                                             # warm-up: the mode switch dropped the cached plan
    torch.cuda.synchronize()
    t0 = time.time()
    rows = []
    for j, data in enumerate(loader):
        i, rn, size = st.comp_rho(data)
        rows.append([st.rho, st.norm])
    torch.cuda.synchronize()
    t_seq = time.time() - t0
    rows = np.array(rows)
    err = float(np.abs(stats[:, 1:3] - rows).max() / np.abs(rows).max())
    print("%s: %d minibatches, world %d: sharded sweep %.3f s, one GPU %.3f s (x%.2f), max rel diff %.2e, weighted rho %.6f" % (
        kind, nb, world, t_sharded, t_seq, t_seq / t_sharded, err, avg[0]))
    assert err < 1e-5
dist.barrier()
dist.destroy_process_group()
