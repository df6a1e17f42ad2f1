"""Worker of tests/test_gpu_multirank.py: one rank of a torchrun launch.  Runs `rounds` coloured RBCD rounds of the
k-agent split with the agents spread over the ranks (public poses by NCCL all-gather) and writes this rank's iterates."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ds, k, rounds, out_dir, accel = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
    conc = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    import torch
    import torch.distributed as dist
    from dpo_b200 import posegraph as pg
    from dpo_b200.agent import DistributedPGO
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    edges, n = pg.read_g2o_file(os.path.join(ROOT, "data", ds + ".g2o"))
    run = DistributedPGO(edges, n, k, r=5, schedule="coloured", rank=rank, world=world, device=local, dist=dist,
                         acceleration=bool(accel), concurrent=bool(conc))
    costs = []
    for _ in range(rounds):
        st = run.step(evaluate=True)
        costs.append((st.cost, st.gradnorm))
    for a in run.local_ids:
        np.save(os.path.join(out_dir, f"X_{a}.npy"), run.agents[a].mProblem.download_X())
    if rank == 0:
        np.save(os.path.join(out_dir, "trace.npy"), np.array(costs))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
