"""Overlapped-schedule tuning on N GPUs (under torchrun): one sharded register, the inverse QFT timed
for every (global SMs, slices, global run bits) setting.  One JSON line per setting on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/mg_tune.py [n] [steps]
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import quantumcomputer_b200 as q  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    p = int(math.log2(world))
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [30 + p]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    for n in sizes:
        ids = [q.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        reg = q.Register(n, 0, device=local_rank, rank=rank, world_size=world, comm_id=ids[0])
        reg.fill_synthetic(1234)
        reg.scale(1.0 / math.sqrt(reg.norm2()))
        for slices in (8, 4, 2):
            for sms in (16, 24, 32, 48, 64):
                reg.set_option(q.OPT_OVERLAP_SLICES, slices)
                reg.set_option(q.OPT_GLOBAL_SMS, sms)
                for _ in range(2):
                    reg.inverse_QFT()
                reg.synchronize()
                dist.barrier()
                reg.set_option(q.OPT_PROFILE, 1)
                reg.profile_reset()
                reg.timer_start()
                for _ in range(steps):
                    reg.inverse_QFT()
                ms = reg.timer_stop()
                prof = reg.profile()
                reg.set_option(q.OPT_PROFILE, 0)
                t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if rank == 0:
                    gs, ts = prof["global_sweep"], prof["tile_sweep"]
                    print(json.dumps({"n": n, "gpus": world, "slices": slices, "global_sms": sms,
                                      "ms_per_iqft": float(t.item()) / steps,
                                      "global_sweep_GBps": gs[2] / (gs[1] * 1e-3) / 1e9 if gs[1] > 0 else None,
                                      "global_sweep_ms": gs[1] / steps,
                                      "local_sweeps_GBps": ts[2] / (ts[1] * 1e-3) / 1e9 if ts[1] > 0 else None,
                                      "local_sweeps_ms": ts[1] / steps}), flush=True)
        nrm = reg.norm2()
        if rank == 0:
            print(json.dumps({"n": n, "norm_after_all": nrm}), flush=True)
        reg.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
