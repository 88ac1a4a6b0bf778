"""torchrun worker of tests/test_dist_nccl_gpu.py: one slab rank per GPU over the NCCL transport; rank 0 compares the
gathered result with a single-device run of the same frames (bit-identical, see tests/test_dist_gpu.py)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    import torch.distributed as dist
    from pbf_sph_b200 import Solver, capi, scenes
    from pbf_sph_b200.dist import SlabRank, shard

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(SlabRank.unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    p, xs = scenes.two_cubes(20000, 4)
    sr = SlabRank(scenes.H, local, rank, world, idt.cpu().numpy().tobytes())
    sr.set_replan(2)
    sr.upload(shard(xs, rank, world))
    for f in range(frames):
        sr.step(scenes.apply_motion(p, f))
    # ... and two more frames with the marching-cubes surface: every rank fills its lattice points into rank 0's arena
    meshes = []
    for f in range(frames, frames + 2):
        pf = scenes.apply_motion(p, f)
        pf.surface_enabled = 1
        sr.step(pf)
        if rank == 0:
            meshes.append(sr.s.mesh())
    mine = sr.download()
    st = sr.stats()
    parts = [None] * world
    dist.all_gather_object(parts, (mine.tobytes(), st))
    ok = True
    if rank == 0:
        got = np.concatenate([np.frombuffer(b, dtype=capi.PARTICLE) for b, _ in parts])
        with Solver(scenes.H, local) as s:
            s.upload(xs)
            for f in range(frames):
                s.step(scenes.apply_motion(p, f))
            ref_meshes = []
            for f in range(frames, frames + 2):
                pf = scenes.apply_motion(p, f)
                pf.surface_enabled = 1
                s.step(pf)
                ref_meshes.append(s.mesh())
            ref = s.download()
        ok = (len(ref) == len(got) and np.array_equal(ref["id"], got["id"])
              and all(np.array_equal(ref[k].view(np.uint32), got[k].view(np.uint32)) for k in ("position", "velocity", "colour")))
        ok = ok and sum(s_["ghosts"] for _, s_ in parts) > 0 and all(s_["owned"] > 0 for _, s_ in parts)
        for a, b in zip(ref_meshes, meshes):  # the slab mesh is the single-device mesh, bit for bit
            ok = ok and len(a.vs) > 0 and len(a.vs) == len(b.vs) and all(
                np.array_equal(getattr(a, k).view(np.uint32), getattr(b, k).view(np.uint32)) for k in ("vs", "ns", "cs"))
        print("NCCL_SLAB_OK" if ok else "NCCL_SLAB_MISMATCH", [s_["owned"] for _, s_ in parts], flush=True)
    sr.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
