"""Batched, sharded dataset synthesis: the throughput counterpart of ``datagen.generate_data``.

The reference generates its 100 k-sample training set with one sequential process
(``datagen/generate.py:56-164``: mesh a plate, then one ``FEAnalysis`` per condition, redrawing a
condition whose solve fails, ``:112-124``).  Here the same dataset tree is produced by a three-stage
pipeline per GPU:

  worker processes  plate geometry + mesh + a stream of candidate condition dicts        (host, CPU)
  main thread       N plates -> CUDA batches: region selection / Dirichlet mask / load and the
                    well-posedness classifier on the device (fea_batch_create_from_conditions),
                    assemble, solve, rasterise u and the region flags
  writer threads    PNG encoding and text files                                            (host, CPU)

Which conditions a plate gets is a function of the plate alone: candidate k of plate p is the k-th
draw of the sampler seeded with ``seed + p``; the plate keeps, in draw order, the first
``conditions_per_plate`` candidates that are well posed (SURVEY A-19; with ``well_posed=False``:
that have no empty matrix row, the reference's own NaN criterion) AND whose solve converges -- the
batched form of the reference's redraw loop.  Nothing that is not a converged solution is ever
written.  Plates are dealt to ranks by ``sharding.plate_shard`` (round-robin, no communication) and
written to ``data_dir/<p+1>/``, so the output of a plate does not depend on the number of ranks,
the batch size or the worker count.

Directory tree (what ``model/diffusion.py::FEADataset`` reads, reference :134-243, 359-378):
  <plate>/input.png, <plate>/outline.png,
  <plate>/<cond>/{outputs_displacement_x.png, outputs_displacement_y.png, regions_<Region>.png,
                  magnitudes.txt, materials.txt, ranges.txt [, domain.<k>.vtk, regions.vtk]
                  [, outputs_{stress,strain}_{x,y}.png]}
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
from .plates import condition_kwargs, make_plate
from .sharding import plate_shard
from .solver import Context, PackedConditions

STATUS_TEXT = {0: "converged", 1: "max_iter", 2: "breakdown", 3: "empty_row", 4: "stagnated"}
FIRST_DRAWS = 3      # candidates drawn up front = FIRST_DRAWS * conditions_per_plate
MAX_DRAWS = 400      # per plate, like workload.plate_conditions


def _plate_job(args):
    """(worker process) mesh of one plate + candidate conditions ``skip .. skip + count - 1`` of its
    sampler stream (one ``sample_conditions`` call per candidate, so that the stream can be resumed)."""
    plate, seed, image_size, mesh_size, skip, count, region_method = args
    try:   # one thread per worker process: the pool already uses every core (scikit-learn / BLAS would each take
        from threadpoolctl import threadpool_limits   # all of them: measured 42 s instead of 4 s of host time per plate)
        with threadpool_limits(1):
            return _plate_job_1(plate, seed, image_size, mesh_size, skip, count, region_method)
    except ImportError:
        return _plate_job_1(plate, seed, image_size, mesh_size, skip, count, region_method)


def _plate_job_1(plate, seed, image_size, mesh_size, skip, count, region_method):
    t0 = time.perf_counter()
    gen, ptags, ltags = make_plate(seed + plate, mesh_size, region_method=region_method)
    coors, conn = gen.mesh
    cands = []
    for k in range(skip + count):
        kw = condition_kwargs(gen.sample_conditions(ptags, ltags, 1)[0])
        if k >= skip:
            cands.append(kw)
    bbox = (float(coors[:, 0].min()), float(coors[:, 1].min()), float(coors[:, 0].max()), float(coors[:, 1].max()))
    window, bounds = imaging.plate_window(bbox, image_size)
    return dict(plate=plate, coors=coors, conn=conn, cands=cands, first=skip, window=window, bounds=bounds,
                affine=imaging.crop_affine(bbox, window, bounds), size=bounds[2] - bounds[0], bbox=bbox,
                host_s=time.perf_counter() - t0)


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


def text_lines(kw: Dict, region_count: np.ndarray):
    """magnitudes.txt / materials.txt of one condition (reference fea_analysis.py:87-91, 108-115,
    278-282; SURVEY A-17): an edge force is listed per vertex, F / max(#region vertices, 1)."""
    vf = kw.get("force_vertex_tags_magnitudes") or ()
    ef = kw.get("force_edges_tags_magnitudes") or ()
    mags = ["VertexForce%d:%s" % (i, str(m)) for i, (_, m) in enumerate(vf)]
    for i, (_, m) in enumerate(ef):
        cnt = max(int(region_count[len(vf) + i]), 1)
        mags.append("EdgeForce%d:%s" % (i, str(tuple(c / cnt for c in m))))
    mats = kw.get("material_properties_to_vertices")
    mat_lines = ["MaterialRegion%d:%s" % (i, str(key)) for i, key in enumerate(mats)] if mats is not None else []
    return "".join(l + "\n" for l in mags), "".join(l + "\n" for l in mat_lines)


def _write_plate(data_dir, job, res, num_steps, save_meshes, save):
    """(writer thread) all files of one plate."""
    plate_dir = os.path.join(data_dir, str(job["plate"] + 1))
    os.makedirs(plate_dir, exist_ok=True)
    _rgb(res["input"]).save(os.path.join(plate_dir, "input.png"))
    _rgb(_outline_image(job["window"], job["bbox"])).save(os.path.join(plate_dir, "outline.png"))
    times = np.linspace(0.0, 1.0, num_steps)
    for ci, r in enumerate(res["conds"]):
        cdir = os.path.join(plate_dir, str(ci + 1))
        os.makedirs(cdir, exist_ok=True)
        for f in os.listdir(cdir):                      # idempotent re-runs (--start_plate resume)
            os.remove(os.path.join(cdir, f))
        if save["displacement"]:
            for c, name in enumerate(("displacement_x", "displacement_y")):
                _rgb(r["images"][c]).save(os.path.join(cdir, "outputs_%s.png" % name))
        for name, img in r.get("cell_images", {}).items():
            _rgb(img).save(os.path.join(cdir, "outputs_%s.png" % name))
        for nm, img in zip(r["names"], r["regions"]):
            _rgb(img).save(os.path.join(cdir, "regions_%s.png" % nm))
        with open(os.path.join(cdir, "magnitudes.txt"), "w") as f:
            f.write(r["magnitudes"])
        with open(os.path.join(cdir, "materials.txt"), "w") as f:
            f.write(r["materials"])
        kinds = []                                      # reference order: displacement, stress, strain (:539-558)
        if save["displacement"]:
            kinds += [("displacement_x", r["ranges"][0:2]), ("displacement_y", r["ranges"][2:4])]
        for name in ("stress_x", "stress_y", "strain_x", "strain_y"):
            if name in r.get("cell_ranges", {}):
                kinds.append((name, r["cell_ranges"][name]))
        with open(os.path.join(cdir, "ranges.txt"), "w") as f:
            for k in range(1, num_steps):
                for name, (lo, hi) in kinds:
                    a, b = float(times[k] * lo), float(times[k] * hi)
                    f.write("%s_%d:%s\n" % (name, k, str((min(a, b), max(a, b)))))
        if save_meshes:
            groups = np.zeros(len(job["coors"]), np.int64)
            mat = np.zeros(len(r["conn"]), np.int64)
            for k, t in enumerate(times):
                write_vtk(os.path.join(cdir, domain_filename(k, num_steps)), job["coors"], r["conn"],
                          point_data={"u": t * r["u"], "node_groups": groups},
                          cell_data={"cauchy_strain": t * r["strain"], "cauchy_stress": t * r["stress"], "mat_id": mat})
            write_vtk(os.path.join(cdir, "regions.vtk"), job["coors"], r["conn"],
                      point_data=dict([("Omega", np.ones(len(job["coors"])))] +
                                      [(nm, fl.astype(np.float64)) for nm, fl in zip(r["names"], r["flags"])]),
                      cell_data={"mat_id": mat})
    return len(res["conds"])


def generate_dataset(data_dir: str, num_plates: int, conditions_per_plate: int = 4, image_size: int = 64,
                     num_steps: int = 11, mesh_size: float = 1e-2, seed: int = 0, rank: int = 0, world: int = 1,
                     device: Optional[int] = None, plates_per_batch: int = 50, workers: Optional[int] = None,
                     writer_threads: int = 8, save_meshes: bool = False, well_posed: bool = True,
                     start_plate: int = 0, rtol: float = 1e-10, max_iter: int = 50000,
                     save_displacement: bool = True, save_stress: bool = False, save_strain: bool = False,
                     progress: Optional[Callable[[int, int], None]] = None, region_method: str = "reference") -> Dict:
    """Generates the plates ``start_plate .. num_plates-1`` owned by ``rank`` of ``world`` into
    ``data_dir``.  Returns throughput statistics.  ``region_method``: how the condition sampler draws the
    material regions -- ``"reference"`` = the reference's KMeans / agglomerative methods (scikit-learn,
    ``mesh_generator.py:319-385``), ``"lloyd"`` = the fast stand-in of the benchmark workloads."""
    assert num_steps > 1, "Must have at least 2 steps per condition."
    os.makedirs(data_dir, exist_ok=True)
    mine = plate_shard(num_plates - start_plate, rank, world, start=start_plate)
    workers = workers or max(1, (os.cpu_count() or 2) // max(world, 1) - 1)
    t1 = float(np.linspace(0.0, 1.0, num_steps)[1])
    save = dict(displacement=save_displacement, stress=save_stress, strain=save_strain)
    want_cells = save_meshes or save_stress or save_strain
    stats = dict(plates=0, samples=0, candidates=0, rejected_ill_posed=0, redrawn_not_converged=0, extra_draw_jobs=0,
                 gpu_s=0.0, host_gen_s=0.0, batches=0)
    t_start = time.perf_counter()
    pending = []

    def classify(jobs: List[dict]):
        """usable[j][k]: candidate k of plate j may be solved (device classifier on every candidate)."""
        meshes = [(j["coors"], j["conn"]) for j in jobs]
        samples = [(ji, kw) for ji, j in enumerate(jobs) for kw in j["cands"][j["classified"]:]]
        if not samples:
            return
        t_c0 = time.perf_counter()
        with ctx.create_batch_from_conditions(PackedConditions(meshes, samples)) as b:
            b.assemble()
            fl, em = b.classify()
        stats["classify_s"] = stats.get("classify_s", 0.0) + time.perf_counter() - t_c0
        ok = (em == 0) & ((fl == 0) | (not well_posed))
        stats["candidates"] += len(samples)
        stats["rejected_ill_posed"] += int((~ok).sum())
        o = 0
        for j in jobs:
            m = len(j["cands"]) - j["classified"]
            j["usable"].extend(bool(x) for x in ok[o:o + m])
            j["classified"] = len(j["cands"])
            o += m

    def solve(jobs: List[dict], picks: List[tuple]):
        """Solves the picked (plate index, candidate index) pairs; returns one result dict per pick."""
        t_s0 = time.perf_counter()
        meshes = [(j["coors"], j["conn"]) for j in jobs]
        samples = [(ji, jobs[ji]["cands"][k]) for ji, k in picks]
        pc = PackedConditions(meshes, samples)
        sz = max(j["size"] for j in jobs)
        affine = np.stack([jobs[ji]["affine"] for ji, _ in picks])
        # the plate mask (input.png) rides along with the first pick of every plate
        seen, mask = set(), np.zeros(len(picks), np.uint8)
        for i, (ji, _) in enumerate(picks):
            if ji not in seen:
                seen.add(ji)
                mask[i] = 1
        with ctx.create_batch_from_conditions(pc) as b:
            b.assemble().solve(rtol, max_iter).rasterize(sz, affine, t1)
            res = b.download(images=True)
            dev = b.setup(flags=save_meshes, counts_only=not save_meshes)
            region_imgs = b.rasterize_regions(mask)
            conn, _ = b.conn() if save_meshes else (None, None)
            strain, stress = b.cell_strain_stress(0) if want_cells else (None, None)
            cell_imgs, cell_rng = {}, {}
            if save_stress or save_strain:
                # per-cell scalar renders (fea_analysis.py:539-558): component x / y of stress and strain at step 1
                names_cf = (["stress_x", "stress_y"] if save_stress else []) + (["strain_x", "strain_y"] if save_strain else [])
                cell_imgs, cell_rng = b.rasterize_cell_components(0, names_cf, t1)
        t_s1 = time.perf_counter()
        stats["solve_gpu_s"] = stats.get("solve_gpu_s", 0.0) + t_s1 - t_s0
        us = pc.split_vertices(res.u)
        out = []
        for i, (ji, k) in enumerate(picks):
            j = jobs[ji]
            n = j["size"]
            reg = region_imgs[i][:, :n, :n]
            d = dict(status=int(res.status[i]), iters=int(res.iters[i]), relres=float(res.relres[i]))
            if d["status"] == SAMPLE_CONVERGED:
                mags, mats = text_lines(j["cands"][k], dev.region_count[i])
                d.update(images=res.images[i][:, :n, :n].copy(), ranges=res.ranges[i].copy(), names=pc.names[i],
                         regions=reg[:len(pc.names[i])].copy(), magnitudes=mags, materials=mats)
                if mask[i]:
                    d["input"] = reg[-1].copy()
                c0, c1 = pc.cell_off[i], pc.cell_off[i + 1]
                if save_meshes:
                    d.update(u=us[i].copy(), conn=conn[c0:c1].copy(), flags=dev.region_flags[i].copy())
                if want_cells:
                    d.update(strain=strain[c0:c1].copy(), stress=stress[c0:c1].copy())
                if cell_imgs:
                    d["cell_images"] = {nm: im[i][:n, :n].copy() for nm, im in cell_imgs.items()}
                    d["cell_ranges"] = {nm: (float(r[i, 0]), float(r[i, 1])) for nm, r in cell_rng.items()}
            elif mask[i]:
                d["input"] = reg[-1].copy()
            out.append(d)
        stats["solve_unpack_s"] = stats.get("solve_unpack_s", 0.0) + time.perf_counter() - t_s1
        return out

    def run_batch(jobs: List[dict], writers: ThreadPoolExecutor, pool):
        t0 = time.perf_counter()
        for j in jobs:
            j.update(classified=0, usable=[], cursor=0, done=[], input=None)
        need = list(range(len(jobs)))
        while need:
            classify(jobs)
            picks = []
            for ji in need:
                j = jobs[ji]
                while len(j["done"]) + sum(1 for p in picks if p[0] == ji) < conditions_per_plate and j["cursor"] < len(j["cands"]):
                    if j["usable"][j["cursor"]]:
                        picks.append((ji, j["cursor"]))
                    j["cursor"] += 1
            if picks:
                for (ji, k), d in zip(picks, solve(jobs, picks)):
                    j = jobs[ji]
                    if j["input"] is None and "input" in d:
                        j["input"] = d["input"]
                    if d["status"] == SAMPLE_CONVERGED:
                        j["done"].append((k, d))
                    else:                                  # the reference redraws a failed condition (generate.py:112-124)
                        stats["redrawn_not_converged"] += 1
            need = [ji for ji in need if len(jobs[ji]["done"]) < conditions_per_plate]
            short = [ji for ji in need if jobs[ji]["cursor"] >= len(jobs[ji]["cands"])]
            if short:                                      # candidate stream exhausted: resume it in the workers
                args = []
                for ji in short:
                    j = jobs[ji]
                    if len(j["cands"]) >= MAX_DRAWS:
                        raise RuntimeError("plate %d: no usable condition in %d draws" % (j["plate"], MAX_DRAWS))
                    args.append((j["plate"], seed, image_size, mesh_size, len(j["cands"]), 2 * FIRST_DRAWS * conditions_per_plate,
                                 region_method))
                stats["extra_draw_jobs"] += len(args)
                for ji, more in zip(short, pool.map(_plate_job, args)):
                    jobs[ji]["cands"].extend(more["cands"])
        stats["gpu_s"] += time.perf_counter() - t0
        for j in jobs:
            j["done"].sort(key=lambda kd: kd[0])           # draw order, whichever round solved it
            out = dict(input=j["input"], conds=[d for _, d in j["done"][:conditions_per_plate]])
            pending.append(writers.submit(_write_plate, data_dir, j, out, num_steps, save_meshes, save))
        stats["batches"] += 1

    # more than half of the sampler's draws are ill-posed: with 3 x conditions_per_plate candidates one
    # plate in seven runs out and is resumed in the workers while the CUDA thread waits.  The stand-in region
    # method is cheap enough to draw 5 x up front; the candidate STREAM of a plate, hence the dataset, is the same
    first = (5 if region_method == "lloyd" else FIRST_DRAWS) * conditions_per_plate
    jobs_args = [(p, seed, image_size, mesh_size, 0, first, region_method) for p in mine]
    ctx = None
    with ThreadPoolExecutor(max_workers=writer_threads) as writers:
        with mp.get_context("fork").Pool(workers) as pool:   # forked BEFORE the CUDA context exists
            ctx = Context(rank if device is None else device)
            batch: List[dict] = []
            for job in pool.imap(_plate_job, jobs_args, chunksize=1):
                stats["host_gen_s"] += job["host_s"]
                batch.append(job)
                if len(batch) >= plates_per_batch:
                    run_batch(batch, writers, pool)
                    stats["plates"] += len(batch)
                    batch = []
                    if progress is not None:
                        progress(stats["plates"], len(mine))
            if batch:
                run_batch(batch, writers, pool)
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
