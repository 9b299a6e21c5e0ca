"""Cross-process strip-mode check (CUDA IPC + NVLink peer loads + flag barrier), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py REBLUR_DIFFUSE_SPECULAR 640 360 4

Every rank denoises its strip; rank 0 also runs the whole frame in a single-GPU context and compares the gathered
strips with it bit for bit.  Prints one JSON line on rank 0 and exits non-zero on a mismatch."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    os.environ["NRD_B200_FORCE_STRIP_KERNELS"] = "1"  # rank 0's full-frame reference runs the strip build of the kernels
    import torch
    import torch.distributed as dist
    from raytracingdenoiser_b200 import harness, nrd, scene, strips
    den_name, w, h, frames = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    den = getattr(nrd.Denoiser, den_name)
    mode = harness.radiance_mode(den)
    part = strips.StripDenoiser(den, w, h, rank, world, device=local)
    part.connect()
    full = harness.GpuDenoiser(den, w, h, device=local) if rank == 0 else None
    sc = scene.Scene(w, h, device="cuda:%d" % local)
    for f in range(frames):
        fr = sc.frame(f, mode)
        cs = harness.make_common_settings(fr, w, h, f)
        part.set_inputs(fr)
        part.denoise(cs)
        if full is not None:
            full.set_inputs(fr)
            full.denoise(cs)
    part.synchronize()
    outs = part.read_outputs()
    torch.cuda.synchronize()
    ok, report = True, {}
    for name in sorted(outs):
        t = outs[name].contiguous()
        # strips are uniform except the last one: pad to the strip height for the gather
        pad = torch.zeros((part.strip_height,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        gathered = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, gathered, dst=0)
        if rank == 0:
            got = torch.cat(gathered, dim=0)[:h]
            ref = full.outputs()[name]
            same = got.view(torch.uint8) == ref.view(torch.uint8)
            report[name] = float(same.float().mean())
            ok = ok and bool(same.all())
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0:
        print(json.dumps({"check": "strips_vs_full_frame", "denoiser": den_name, "size": [w, h], "frames": frames, "world": world, "identical": report, "ok": ok}))
    dist.barrier()
    part.destroy()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
