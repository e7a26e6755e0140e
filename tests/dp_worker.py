"""Worker of tests/test_gpu_multi.py: one rank of a 2-GPU data-parallel regression run (launched by torchrun)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_path):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from deeppicarditeration_b200.solution import construct_mlp
    from deeppicarditeration_b200.trainer import RegressionTrainer

    nx, n, gb, epochs = 100, 4096, 512, 3
    g = torch.Generator().manual_seed(7)
    tx = torch.cat([torch.rand(n, 1, generator=g) * 0.98 + 0.01, torch.randn(n, nx, generator=g)], 1).cuda()
    y = (torch.randn(n, 1 + nx, generator=g) * 0.3).cuda()

    def model():
        torch.manual_seed(1)
        return construct_mlp(1 + nx, 1, [128] * 4, ["ELU"] * 4, None).float().cuda()

    # data parallel: every global mini-batch of 512 rows is split into contiguous halves (rank order)
    lb = gb // world
    idx = torch.cat([torch.arange(b + rank * lb, b + (rank + 1) * lb) for b in range(0, n, gb)]).cuda()
    m_dp = model()
    tr = RegressionTrainer(m_dp, fixed_weight=0.7, loss_beta=0.3, max_batch=lb)
    assert tr.dp_world == world
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tr.fit(tx[idx], y[idx], lb, 1)  # warm-up: captures the graph
    torch.cuda.synchronize()
    dist.barrier()
    ev0.record()
    tr.fit(tx[idx], y[idx], lb, epochs - 1)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    flat = torch.cat([p.detach().reshape(-1) for p in m_dp.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    bitwise = all(torch.equal(gathered[0], t) for t in gathered[1:])
    res = {"rank": rank, "replicas_bitwise_identical": bool(bitwise), "ms_per_step": ms / ((epochs - 1) * (n // gb))}
    if rank == 0:
        # the 1-GPU trajectory at the global batch: same data order, same steps
        m_1 = model()
        tr1 = RegressionTrainer(m_1, fixed_weight=0.7, loss_beta=0.3, max_batch=gb, data_parallel=False)
        tr1.fit(tx, y, gb, epochs)
        torch.cuda.synchronize()
        flat1 = torch.cat([p.detach().reshape(-1) for p in m_1.parameters()])
        res["rel_diff_vs_single_gpu"] = float((flat - flat1).norm() / flat1.norm())
        res["max_abs_diff_vs_single_gpu"] = float((flat - flat1).abs().max())
        tr1.close()
    tr.close()
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(allres, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
