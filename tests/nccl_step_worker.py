"""Worker of tests/test_gpu_dist_nccl.py (one process per GPU, launched by torch.distributed.run): two protected
IPR-DCGAN steps, global batch split with torch.chunk, gradients all-reduced over NCCL, against the oracle's
model of the reference's nn.DataParallel semantics (oracle.ShardedStepOracle: per-chunk BatchNorm, averaged
gradients, losses averaged over the whole batch)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    use_graph = sys.argv[1] == "graph"
    global_batch = int(sys.argv[2])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from ipr_gan_b200 import dist as idist
    from ipr_gan_b200.trainer import ProtectedDCGANTrainer
    tr = ProtectedDCGANTrainer(global_batch // world, dev, use_graph=False)
    init_g = {k: v.detach().cpu().clone() for k, v in tr.model.G.module.state_dict().items()}
    init_d = {k: v.detach().cpu().clone() for k, v in tr.model.D.module.state_dict().items()}
    from oracle import ipr_oracle as orc
    batches = [orc.synth_step_inputs(global_batch, seed=77 + i) for i in range(2)]
    if use_graph:
        # capture() warms up with real steps, which would move the weights: capture on a throw-away copy of the state
        tr.use_graph = True
        tr.set_inputs(idist.shard(batches[0][0]), idist.shard(batches[0][1]))
        tr.capture(warmup=1)
        tr.model.G.module.load_state_dict(init_g)
        tr.model.D.module.load_state_dict(init_d)
        for opt in (tr.model.optG, tr.model.optD):
            opt._m.zero_(), opt._v.zero_(), opt._step.zero_()
        from ipr_gan_b200 import engine
        engine.reset_caches()
    got = []
    for real, z in batches:
        m = tr.step_from_host(idist.shard(real), idist.shard(z))
        allm = [None] * world
        dist.all_gather_object(allm, m)
        got.append(allm)
    torch.cuda.synchronize()
    # replicas stay bit-identical: same reduced gradients applied to the same weights
    flat = torch.cat([p.detach().reshape(-1) for p in list(tr.model.G.parameters()) + list(tr.model.D.parameters())])
    others = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(others, flat)
    same = all(torch.equal(others[0], o) for o in others)
    out = {"ok": False}
    if rank == 0:
        mark = os.path.join(ROOT, "ipr_gan_b200", "assets", "watermark_a.png")
        fg, bg = orc.load_watermark(mark, 16, True, True)

        def replica():
            G, D = orc.make_generator(), orc.make_discriminator()
            G.load_state_dict(init_g)
            D.load_state_dict(init_d)
            return orc.DCGANStepOracle(G, D, orc.transform_dist, lambda y: orc.paste_patch(y, fg, bg, "tl", 16))
        ref = orc.ShardedStepOracle(replica, world)
        worst, detail = 0.0, []
        for (real, z), allm in zip(batches, got):
            ref.step(real, z)
            want = ref.metrics()
            mean = {k: sum(m[k] for m in allm) / world for k in want}
            for k in want:
                err = abs(mean[k] - want[k]) / max(1.0, abs(want[k]))
                worst = max(worst, err)
                detail.append((k, mean[k], want[k]))
            # after the cross-rank reduction every rank reports the global value
            spread = max(abs(m[k] - allm[0][k]) for m in allm for k in want)
        pmax = 0.0
        for p, q in zip(tr.model.G.module.parameters(), ref.replicas[0].G.parameters()):
            pmax = max(pmax, float((p.detach().cpu() - q.detach()).abs().max()))
        out = {"ok": True, "replicas_identical": bool(same), "worst_metric_rel": worst, "rank_spread": spread,
               "param_max_abs": pmax, "detail": detail, "world": world, "graph": use_graph}
        print("RESULT " + json.dumps(out))
    sys.stdout.flush()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
