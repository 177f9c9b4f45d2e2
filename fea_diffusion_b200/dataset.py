"""Batched, sharded dataset synthesis: the throughput counterpart of ``datagen.generate_data``.

The reference generates its 100 k-sample training set with one sequential process
(``datagen/generate.py:56-164``: mesh a plate, then one ``FEAnalysis`` per condition).  Here the same
dataset tree is produced by a three-stage pipeline per GPU:

  worker processes  plate geometry + mesh + well-posed conditions + region selection   (host, CPU)
  main thread       pack N plates -> one CUDA batch: assemble, solve, rasterise u and region flags
  writer threads    PNG encoding and text files                                            (host, CPU)

Plates are dealt to ranks by ``sharding.plate_shard`` (round-robin, no communication); plate p is
always generated from seed ``seed + p`` and written to ``data_dir/<p+1>/``, so the output of a
plate does not depend on the number of ranks, the batch size or the worker count.

Directory tree (what ``model/diffusion.py::FEADataset`` reads, reference :134-243, 359-378):
  <plate>/input.png, <plate>/outline.png,
  <plate>/<cond>/{outputs_displacement_x.png, outputs_displacement_y.png, regions_<Region>.png,
                  magnitudes.txt, materials.txt, ranges.txt [, domain.<k>.vtk, regions.vtk]}
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, List, Optional

import numpy as np
from PIL import Image

from . import imaging
from ._capi import SAMPLE_CONVERGED
from .datagen.vtk_io import domain_filename, write_vtk
from .sharding import plate_shard
from .solver import Context, pack
from .workload import plate_conditions

STATUS_TEXT = {0: "converged", 1: "max_iter", 2: "breakdown", 3: "empty_row", 4: "stagnated"}


def _plate_job(args):
    """(worker process) everything the GPU stage needs for one plate."""
    plate, seed, conditions, image_size, mesh_size, well_posed = args
    t0 = time.perf_counter()
    items, rejected = plate_conditions(seed + plate, conditions, image_size, mesh_size, well_posed)
    n_v = len(items[0].setup.coors)
    conds = []
    for it in items:
        names = list(it.setup.regions)
        flags = np.zeros((len(names), n_v), np.uint8)
        for i, nm in enumerate(names):
            flags[i, it.setup.regions[nm]] = 1
        conds.append(dict(sample=it.setup.sample, names=names, flags=flags,
                          magnitudes="".join(l + "\n" for l in it.setup.magnitudes_lines),
                          materials="".join(l + "\n" for l in it.setup.materials_lines)))
    it0 = items[0]
    return dict(plate=plate, conds=conds, window=it0.window, bounds=it0.bounds, affine=it0.affine, size=it0.size,
                bbox=it0.setup.bbox(), rejected=rejected, host_s=time.perf_counter() - t0)


def _rgb(gray: np.ndarray) -> Image.Image:
    return Image.fromarray(np.repeat(gray[:, :, None], 3, axis=2))


def _outline_image(window: int, bbox) -> np.ndarray:
    gray = np.full((window, window), 255, np.uint8)
    l, t, r, b = (max(0, min(window - 1, v)) for v in imaging.outline_bounds(window, bbox))
    for w in (0, 1):
        gray[min(t + w, b), l:r + 1] = 186
        gray[max(b - w, t), l:r + 1] = 186
        gray[t:b + 1, min(l + w, r)] = 186
        gray[t:b + 1, max(r - w, l)] = 186
    return gray


def _write_plate(data_dir, job, res, num_steps, save_meshes):
    """(writer thread) all files of one plate."""
    plate_dir = os.path.join(data_dir, str(job["plate"] + 1))
    os.makedirs(plate_dir, exist_ok=True)
    _rgb(res["input"]).save(os.path.join(plate_dir, "input.png"))
    _rgb(_outline_image(job["window"], job["bbox"])).save(os.path.join(plate_dir, "outline.png"))
    times = np.linspace(0.0, 1.0, num_steps)
    for ci, cond in enumerate(job["conds"]):
        cdir = os.path.join(plate_dir, str(ci + 1))
        os.makedirs(cdir, exist_ok=True)
        r = res["conds"][ci]
        for c, name in enumerate(("displacement_x", "displacement_y")):
            _rgb(r["images"][c]).save(os.path.join(cdir, "outputs_%s.png" % name))
        for nm, img in zip(cond["names"], r["regions"]):
            _rgb(img).save(os.path.join(cdir, "regions_%s.png" % nm))
        with open(os.path.join(cdir, "magnitudes.txt"), "w") as f:
            f.write(cond["magnitudes"])
        with open(os.path.join(cdir, "materials.txt"), "w") as f:
            f.write(cond["materials"])
        rg = r["ranges"]
        with open(os.path.join(cdir, "ranges.txt"), "w") as f:
            for k in range(1, num_steps):
                f.write("displacement_x_%d:%s\n" % (k, str((float(times[k] * rg[0]), float(times[k] * rg[1])))))
                f.write("displacement_y_%d:%s\n" % (k, str((float(times[k] * rg[2]), float(times[k] * rg[3])))))
        if r["status"] != SAMPLE_CONVERGED:
            with open(os.path.join(cdir, "status.txt"), "w") as f:
                f.write("%s iterations=%d relres=%.3e\n" % (STATUS_TEXT.get(r["status"], "?"), r["iters"], r["relres"]))
        if save_meshes:
            smp = cond["sample"]
            groups = np.zeros(len(smp.coors), np.int64)
            mat = np.zeros(len(r["conn"]), np.int64)
            for k, t in enumerate(times):
                write_vtk(os.path.join(cdir, domain_filename(k, num_steps)), smp.coors, r["conn"],
                          point_data={"u": t * r["u"], "node_groups": groups},
                          cell_data={"cauchy_strain": t * r["strain"], "cauchy_stress": t * r["stress"], "mat_id": mat})
            write_vtk(os.path.join(cdir, "regions.vtk"), smp.coors, r["conn"],
                      point_data=dict([("Omega", np.ones(len(smp.coors)))] +
                                      [(nm, fl.astype(np.float64)) for nm, fl in zip(cond["names"], cond["flags"])]),
                      cell_data={"mat_id": mat})
    return len(job["conds"])


def generate_dataset(data_dir: str, num_plates: int, conditions_per_plate: int = 4, image_size: int = 64,
                     num_steps: int = 11, mesh_size: float = 1e-2, seed: int = 0, rank: int = 0, world: int = 1,
                     device: Optional[int] = None, plates_per_batch: int = 50, workers: Optional[int] = None,
                     writer_threads: int = 8, save_meshes: bool = False, well_posed: bool = True,
                     start_plate: int = 0, rtol: float = 1e-10, max_iter: int = 50000,
                     progress: Optional[Callable[[int, int], None]] = None) -> Dict:
    """Generates the plates ``start_plate .. num_plates-1`` owned by ``rank`` of ``world`` into
    ``data_dir``.  Returns throughput statistics."""
    assert num_steps > 1, "Must have at least 2 steps per condition."
    os.makedirs(data_dir, exist_ok=True)
    mine = plate_shard(num_plates - start_plate, rank, world, start=start_plate)
    workers = workers or max(1, (os.cpu_count() or 2) // max(world, 1) - 1)
    t1 = float(np.linspace(0.0, 1.0, num_steps)[1])
    stats = dict(plates=0, samples=0, not_converged=0, rejected_draws=0, gpu_s=0.0, host_gen_s=0.0, batches=0)
    t_start = time.perf_counter()
    pending = []

    def run_batch(jobs: List[dict], writers: ThreadPoolExecutor):
        # crop sizes differ by +-1 px between plates: render the batch at the largest size (the affine
        # maps are relative to each crop's origin) and cut every sample's top-left square out of it
        sz = max(j["size"] for j in jobs)
        samples = [c["sample"] for j in jobs for c in j["conds"]]
        affine = np.stack([j["affine"] for j in jobs for _ in j["conds"]])
        packed = pack(samples)
        t0 = time.perf_counter()
        with ctx.create_batch(packed) as b:
            b.assemble().solve(rtol, max_iter).rasterize(sz, affine, t1)
            res = b.download(images=True)
            flags = []
            for j in jobs:
                for ci, c in enumerate(j["conds"]):
                    f = c["flags"]
                    if ci == 0:   # the plate mask (input.png) rides along as one more field
                        f = np.concatenate([f, np.ones((1, f.shape[1]), np.uint8)])
                    flags.append(f)
            region_imgs = b.rasterize_flags(flags)
            conn, _ = b.conn() if save_meshes else (None, None)
            strain, stress = b.cell_strain_stress(0) if save_meshes else (None, None)
        stats["gpu_s"] += time.perf_counter() - t0
        us = packed.split_vertices(res.u)
        k = 0
        for j in jobs:
            out = dict(conds=[])
            n = j["size"]
            for ci, c in enumerate(j["conds"]):
                reg = region_imgs[k][:, :n, :n]
                if ci == 0:
                    out["input"] = reg[-1]
                    reg = reg[:-1]
                d = dict(images=res.images[k][:, :n, :n], regions=reg, ranges=res.ranges[k], status=int(res.status[k]),
                         iters=int(res.iters[k]), relres=float(res.relres[k]))
                if save_meshes:
                    c0, c1 = packed.cell_off[k], packed.cell_off[k + 1]
                    d.update(u=us[k], conn=conn[c0:c1], strain=strain[c0:c1], stress=stress[c0:c1])
                stats["not_converged"] += int(res.status[k] != SAMPLE_CONVERGED)
                out["conds"].append(d)
                k += 1
            pending.append(writers.submit(_write_plate, data_dir, j, out, num_steps, save_meshes))
        stats["batches"] += 1

    jobs_args = [(p, seed, conditions_per_plate, image_size, mesh_size, well_posed) for p in mine]
    ctx = None
    with ThreadPoolExecutor(max_workers=writer_threads) as writers:
        with mp.get_context("fork").Pool(workers) as pool:   # forked BEFORE the CUDA context exists
            ctx = Context(rank if device is None else device)
            batch: List[dict] = []
            for job in pool.imap(_plate_job, jobs_args, chunksize=1):
                stats["host_gen_s"] += job["host_s"]
                stats["rejected_draws"] += job["rejected"]
                batch.append(job)
                if len(batch) >= plates_per_batch:
                    run_batch(batch, writers)
                    stats["plates"] += len(batch)
                    batch = []
                    if progress is not None:
                        progress(stats["plates"], len(mine))
            if batch:
                run_batch(batch, writers)
                stats["plates"] += len(batch)
                if progress is not None:
                    progress(stats["plates"], len(mine))
        for f in pending:
            stats["samples"] += f.result()
    ctx.close()
    stats["wall_s"] = time.perf_counter() - t_start
    stats["samples_per_s"] = stats["samples"] / stats["wall_s"] if stats["wall_s"] > 0 else 0.0
    stats["gpu_samples_per_s"] = stats["samples"] / stats["gpu_s"] if stats["gpu_s"] > 0 else 0.0
    stats.update(rank=rank, world=world, workers=workers)
    return stats
