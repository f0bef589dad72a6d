"""Data-parallel gradient equality on NCCL: N ranks x batch B, bucket all-reduced with ReduceOp.AVG == 1 process x batch N*B.
Launch:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dp_grad_check.py
Every rank also computes the big-batch gradient itself (same seeds), so the check needs no extra communication; rank 0 prints a
JSON line.  The losses are means over the batch, so the mean of the per-rank gradients is the gradient of the concatenated batch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gs_dynamics_b200 import dist as gdist, gnn, gnn_train, workloads as GO


def main():
    rank, local_rank, world = gdist.env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    gdist.init(backend="nccl", device=dev)
    B, n_obj, n_future, nf = 8, 60, 3, 128
    cfg = GO.sloth_cfg(nf)
    funcs = gnn_train.default_loss_funcs({"mse_loss": 1.0, "length_loss": 0.05})

    def grads(batch_cpu):
        model = gnn.DynamicsPredictor(dict(cfg), dev).to(dev).train()
        model.load_state_dict(GO.make_state_dict(cfg, 0, head_scale=0.05))
        batch = {k: v.to(dev) for k, v in batch_cpu.items()}
        batch["Rr"] = gnn.construct_edges_index(batch["state"][:, -1], 0.075, batch["state_mask"], batch["eef_mask"], topk=5, connect_all=True)
        batch["Rs"] = None
        batch["max_nR"] = batch["Rr"].capacity       # same denominator for every batch size
        bucket = gnn_train.GradientBucket(model.parameters())
        bucket.zero()
        loss, _ = gnn_train.unrolled_loss(model, batch, n_future, funcs)
        loss.backward()
        return bucket, float(loss)

    parts = [GO.make_training_batch(B, n_obj, 500 + r, "sloth", n_future, learnable=True) for r in range(world)]
    mine, loss_mine = grads(parts[rank])
    mine.all_reduce_mean()                                            # the collective under test (NCCL AVG)
    big_batch = {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
    big, loss_big = grads(big_batch)
    torch.cuda.synchronize()
    err = float((mine.flat - big.flat).abs().max() / big.flat.abs().max())
    losses = torch.tensor([loss_mine], device=dev, dtype=torch.float64)
    dist.all_reduce(losses)
    ok = err < 1e-5 and abs(float(losses[0]) / world - loss_big) < 1e-6 * abs(loss_big) + 1e-9
    if rank == 0:
        print(json.dumps({"check": "N ranks x B averaged gradient == 1 rank x N*B gradient (NCCL ReduceOp.AVG)", "world": world, "batch_per_rank": B,
                          "max_rel_err": err, "mean_rank_loss": float(losses[0]) / world, "big_batch_loss": loss_big, "ok": bool(ok)}), flush=True)
    gdist.finalize()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
