#!/usr/bin/env python
"""bench.py -- DP cells/s and decoded reads/s of the batched Viterbi decoder (--error-global).

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  A "step" is one pass of the hot path (fill + traceback kernels) over one batch of
synthetic reads per GPU; reads are independent, so ranks shard the batch with no collective
("scaling": "weak": reads per GPU per step are fixed).

  value     whole-job DP cells/s with the packed reads already resident in HBM (CUDA events on
            the launching stream, barrier + synchronize on both sides, max over ranks)
  e2e       the same metric through the C-ABI host-buffer call dnab_viterbi_batch (pinned host
            buffers, H2D of reads and D2H of decoded strings / log-likelihoods inside the timed region)
  roofline  HBM roofline of the dominant kernel (viterbiFillPushKernel): algorithmic bytes
            = 1 B per DP cell + ceil(L/4) + |decoded| + 8 per read / its CUDA-event time
  cpu_baseline  the reference's own CPU decoder (oracle/_ref/dnastore when it was built, else the
            oracle port) on a bounded sample of the same reads, one core, decoded strings compared

`--impl reference` times the reference CPU implementation on all host cores (one process per core,
disjoint reads), same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "cfg2": dict(recipe=("l4c4", "flusher", "mixradar6"), pool="cfg2_flusher_mixradar6_l4c4_204b", length=4,
                 mut=dict(sub_rate=0.01), desc="flusher*mixradar6*dnastore-l4 (46,670 states), ~200-nt reads, 1% substitutions"),
    "cfg1": dict(recipe=("l4c4",), pool="cfg1_l4c4_200b", length=4, mut=dict(sub_rate=0.01),
                 desc="dnastore-l4 (384 states), ~200-nt reads, 1% substitutions"),
    "cfg3": dict(recipe=("l4c4", "sync16", "flusher", "hamming74"), pool="cfg3_s16h74l4c4_92b", length=4,
                 mut=dict(sub_rate=0.01, dup_rate=0.01, max_dup=2, del_rate=0.01, max_del=4),
                 desc="sync16*hamming74*dnastore-l4 (12,361 states), indels"),
    "cfg4": dict(recipe=("l4c4", "water64.1"), pool="cfg4_water64.1_l4c4_64b", length=4,
                 mut=dict(sub_rate=0.01, del_rate=0.01, max_del=4), desc="watermark64.1*dnastore-l4 (7,066 states)"),
    # BASELINE.json configs[4]'s machine (dnastore -l 8, 10,746 states, k = 4), 150-bit payloads, substitutions
    "cfg5": dict(recipe=("l8c4",), pool="cfg5_l8c4_150b", length=8, mut=dict(sub_rate=0.01),
                 desc="dnastore-l8 (10,746 states, k = 4), ~150-bit payloads, 1% substitutions"),
}
METRIC = "viterbi_dp_cells_per_sec"
UNIT = "cells/s"


def make_reads(w, n, seed):
    from benchdata import synth
    pool = synth.load_pool(w["pool"])
    rng = np.random.default_rng(seed)
    base = [pool[i % len(pool)] for i in range(n)]
    mut = w["mut"]
    if set(mut) == {"sub_rate"}:
        return synth.mutate_subs_batch(base, rng, sub_rate=mut["sub_rate"])
    return [synth.mutate(s, rng, **mut) for s in base]


def cells_of(n_states, k, read_len):
    return int(n_states) * int(np.sum(read_len.astype(np.int64) + 1)) * (k + 2)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                    power_w_max=float(max(power)), samples=len(sm))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------
# reference arm / CPU baseline
# ----------------------------------------------------------------------------------------------
def _ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "dnastore")
    return p if os.path.exists(p) and os.access(p, os.X_OK) else None


def _write_machine_json(machine):
    tf = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
    tf.write(machine.to_json())
    tf.close()
    return tf.name


def cpu_reference_run(w, machine, compiled, reads, n_procs):
    """Decode `reads` with the reference CPU implementation on n_procs cores (one process per core,
    disjoint contiguous shards). Returns (seconds, decoded strings, kind)."""
    binary = _ref_binary()
    shards = [reads[i::n_procs] for i in range(n_procs)]
    shards = [s for s in shards if s]
    if binary:
        mjson = _write_machine_json(machine)
        files = []
        for s in shards:
            tf = tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False)
            for i, r in enumerate(s):
                tf.write(f">r{i}\n{r}\n")
            tf.close()
            files.append(tf.name)
        base = [binary, "-v0", "-l", str(w["length"]), "--load-machine", mjson]
        # machine-load-only run, subtracted: the reference re-parses the JSON per process
        t0 = time.perf_counter()
        subprocess.run(base + ["--save-machine", os.devnull], check=True, capture_output=True)
        t_load = time.perf_counter() - t0
        t0 = time.perf_counter()
        procs = [subprocess.Popen(base + ["-V", f, "--error-global", "--raw"] + list(w.get("ref_flags", [])), stdout=subprocess.PIPE, text=True)
                 for f in files]
        outs = [p.communicate()[0] for p in procs]
        dt = time.perf_counter() - t0 - t_load
        for f in files + [mjson]:
            os.unlink(f)
        dec_shards = [o.split("\n")[:len(s)] for o, s in zip(outs, shards)]
        kind = "reference"
    else:
        import dnab_testutil as util
        from concurrent.futures import ThreadPoolExecutor  # ctypes releases the GIL inside the oracle

        def work(shard):
            return [util.oracle_viterbi(compiled, r, want_path=False)["decoded"] for r in shard]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(len(shards)) as ex:
            dec_shards = list(ex.map(work, shards))
        dt = time.perf_counter() - t0
        kind = "port"
    decoded = [None] * len(reads)
    for i, ds in enumerate(dec_shards):
        decoded[i::n_procs] = ds
    return max(dt, 1e-9), decoded, kind


def bench_fwdback(args, w, compiled, dec, rank, local_rank, world, dev, dist, torch, d, util):
    """Forward-backward step: dnab_fwdback_counts_batch over one batch of reads per GPU (host buffers; the
    copies are a few hundred bytes per read). Algorithmic bytes: 16 per DP cell (the forward cell is written
    once and read once by the backward pass)."""
    t = compiled.t
    rps = args.reads_per_step or {"cfg5": 1184, "cfg1": 8192}.get(args.workload, 592)
    n_batches = args.warmup + args.steps
    batches = [make_reads(w, rps, seed=(rank + 1) * 200003 + b) for b in range(n_batches)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for b in range(args.warmup):
        dec.fwdback_counts(batches[b])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    kernel_ms, cells, max_gap = 0.0, 0, 0.0
    t0 = time.perf_counter()
    for b in range(args.warmup, n_batches):
        out = dec.fwdback_counts(batches[b])
        kernel_ms += dec.stats()["last_fill_ms"]  # CUDA events around the kernel, on its stream
        cells += cells_of(t.n_states, t.k, np.array([len(r) for r in batches[b]]))
        max_gap = max(max_gap, float(np.abs(out["loglike_back"] - out["loglike"]).max()))
        assert (out["status"] == 0).all()
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    vals = torch.tensor([kernel_ms, wall_s], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(cells), float(rps * args.steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    kernel_ms_max, wall_max = vals.tolist()
    tot_cells, tot_reads = sums.tolist()
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = 16.0 * cells / (kernel_ms * 1e-3) / 1e9
        cpu = None
        if args.cpu_sample > 0:
            sample = batches[-1][:min(args.cpu_sample, 2)]
            t1 = time.perf_counter()
            ref = [util.oracle_fwdback(compiled, r) for r in sample]
            dt = time.perf_counter() - t1
            for i, o in enumerate(ref):
                assert util.hexf(o["loglike"]) == util.hexf(out["loglike"][i]), "GPU forward log-likelihood differs from the specification"
                np.testing.assert_allclose(out["counts"][i], o["counts"], rtol=1e-9, atol=1e-12)
            cpu_cells = cells_of(t.n_states, t.k, np.array([len(r) for r in sample]))
            cpu = dict(value=cpu_cells / dt, unit=UNIT, cores=1, kind="port", reads_per_sec=len(sample) / dt,
                       sample=f"first {len(sample)} reads of the last timed batch through oracle/forward_oracle.c (a specification: "
                              f"the reference has no machine-lattice forward-backward), results equal to the GPU's; host has {os.cpu_count()} cores")
        line = dict(metric="fwdback_dp_cells_per_sec", value=tot_cells / (kernel_ms_max * 1e-3), unit=UNIT, n_gpus=world,
                    steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * wall_max / args.steps, higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                    reads_per_sec=tot_reads / (kernel_ms_max * 1e-3),
                    config=dict(workload=f"{args.workload}: {w['desc']}, --error-global, -l {w['length']}; forward + backward + "
                                         "posterior counts (parity unpinned: not in the reference)",
                                reads_per_step_per_gpu=rps, n_states=int(t.n_states), k=int(t.k),
                                l2="distinct read batch per step; the forward cells (8 B each) are streamed to HBM and read back",
                                max_abs_loglike_back_minus_forward=max_gap),
                    e2e=dict(value=tot_cells / wall_max, unit=UNIT, h2d_bytes_per_step=int(sum((len(r) + 3) // 4 + 12 for r in batches[-1])),
                             d2h_bytes_per_step=int(rps * (8 * (5 + t.k + 16) + 20)), steps=args.steps,
                             reads_per_sec=tot_reads / wall_max),
                    gpu_launches=int(args.steps * world),
                    roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=None,
                                  kernel="forwardKernel", launches=args.steps, avg_launch_ms=kernel_ms / args.steps,
                                  peak_source=peak_src, algorithmic_bytes_per_launch=16.0 * cells / args.steps),
                    cpu_baseline=cpu, clocks=clocks)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads-per-step", type=int, default=0, help="reads per GPU per step (0 = default for the workload)")
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--tmode", type=int, default=0)
    ap.add_argument("--table-mode", type=int, default=0)
    ap.add_argument("--partition", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], help="decoder option key=value (dnab_decoder_set_option), repeatable")
    ap.add_argument("--cpu-sample", type=int, default=4, help="reads in the single-core CPU baseline sample (0 = skip)")
    ap.add_argument("--no-indel", action="store_true", help="decode with --error-del-open 0 --error-dup-prob 0 (closure degenerates)")
    ap.add_argument("--mode", default="viterbi", choices=["viterbi", "fwdback"],
                    help="fwdback: forward + backward + posterior counts over the machine lattice (SURVEY 8a-12, BASELINE "
                         "configs[4]; not in the reference, CPU baseline = the specification in oracle/forward_oracle.c)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = WORKLOADS[args.workload]
    if args.no_indel:
        # the same workload decoded with --error-del-open 0 --error-dup-prob 0 -- the within-column
        # closure degenerates to one visit per state, an upper bound on what the exact closure costs
        w = dict(w, flags=dict(del_open=0., dup_prob=0.), ref_flags=["--error-del-open", "0", "--error-dup-prob", "0"],
                 desc=w["desc"] + ", --error-del-open 0 --error-dup-prob 0")

    import __graft_entry__ as g
    if rank == 0:
        g.build()
    import dnab_testutil as util
    import dnastore_b200 as d

    if args.impl == "reference":
        if rank != 0:
            return
        machine = util.machine_from_recipe(w["recipe"])
        compiled = machine.compile(d.ErrorFlags(length=w["length"], global_=True, **w.get("flags", {})))
        t = compiled.t
        cores = os.cpu_count() or 1
        per_step = cores  # one read per core per step: a bounded sample of the same workload
        times, cells = [], []
        for step in range(args.warmup + args.steps):
            reads = make_reads(w, per_step, seed=1000 + step)
            dt, _dec, kind = cpu_reference_run(w, machine, compiled, reads, cores)
            if step >= args.warmup:
                times.append(dt)
                cells.append(cells_of(t.n_states, t.k, np.array([len(r) for r in reads])))
        total_t = sum(times)
        value = sum(cells) / total_t
        line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=1e3 * total_t / max(args.steps, 1), higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                    reads_per_sec=per_step * args.steps / total_t,
                    config=dict(workload=f"{args.workload}: {w['desc']}, --error-global, -l {w['length']}",
                                reads_per_step=per_step, n_states=int(t.n_states), k=int(t.k)),
                    cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=kind,
                                      sample=f"{per_step} reads per step (one per core), {args.steps} steps"),
                    e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    dev = torch.device("cuda", local_rank)

    machine = util.machine_from_recipe(w["recipe"])
    compiled = machine.compile(d.ErrorFlags(length=w["length"], global_=True, **w.get("flags", {})))
    t = compiled.t
    dec = d.Decoder(compiled, device=local_rank)
    if args.cluster or args.threads or args.tmode or args.table_mode or args.partition:
        dec.configure(args.cluster, args.threads, args.tmode, args.table_mode, args.partition)
    for kv in args.opt:
        key, value = kv.split("=")
        dec.set_option(key, int(value))
    if args.mode == "fwdback":
        return bench_fwdback(args, w, compiled, dec, rank, local_rank, world, dev, dist, torch, d, util)
    binfo = dec.batch_info()
    info = dec.info() if not binfo["enabled"] else None
    default_rps = {"cfg2": 960, "cfg1": 65536, "cfg3": 4096, "cfg4": 8192, "cfg5": 4096}[args.workload]
    rps = args.reads_per_step or default_rps

    # distinct batch per step, generated before timing and resident in HBM
    n_batches = args.warmup + args.steps
    batches = []
    for b in range(n_batches):
        reads = make_reads(w, rps, seed=(rank + 1) * 100003 + b)
        packed, byte_off, read_len = d.pack_reads(reads)
        batches.append((reads if b == n_batches - 1 else None, packed, byte_off, read_len))
    max_len = max(int(b[3].max()) for b in batches)
    stride = 2 * max_len + 64  # decoded strings are ~1 symbol per base for these codes

    def to_dev(a):
        return torch.from_numpy(a).to(dev)

    dev_batches = [(to_dev(p), to_dev(o), to_dev(l)) for (_r, p, o, l) in batches]
    d_ll = torch.zeros(rps, dtype=torch.float64, device=dev)
    d_dec = torch.zeros(rps * stride, dtype=torch.uint8, device=dev)
    d_declen = torch.zeros(rps, dtype=torch.int32, device=dev)
    d_status = torch.zeros(rps, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def run_device(b):
        p, o, l = dev_batches[b]
        dec.viterbi_device(rps, max_len, p.data_ptr(), o.data_ptr(), l.data_ptr(), d_ll.data_ptr(), d_dec.data_ptr(),
                           stride, d_declen.data_ptr(), d_status.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for b in range(args.warmup):
        run_device(b)
    barrier()
    dec.set_timing(True)
    dec.reset_timing()
    launches0 = dec.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for b in range(args.warmup, n_batches):
        run_device(b)
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    st = dec.stats()
    dec.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    gpu_launches = int(st["kernel_launches"] - launches0)
    status_ok = bool((d_status == 0).all().item())
    dec_len_last = d_declen.cpu().numpy().astype(np.int64)

    my_cells = sum(cells_of(t.n_states, t.k, batches[b][3]) for b in range(args.warmup, n_batches))
    my_reads = rps * args.steps

    # ---- e2e: host buffers through dnab_viterbi_batch (pinned), copies inside the timed region --------
    def pin(a):
        tp = torch.from_numpy(a).pin_memory()
        return tp, tp.numpy()
    e2e_steps = max(1, min(args.steps, 2))
    e2e_bufs = []
    for b in range(n_batches - e2e_steps, n_batches):
        _r, p, o, l = batches[b]
        e2e_bufs.append((pin(p), pin(o), pin(l)))
    h_ll = pin(np.zeros(rps, dtype=np.float64))
    h_dec = pin(np.zeros(rps * stride, dtype=np.uint8))
    h_declen = pin(np.zeros(rps, dtype=np.int32))
    h_status = pin(np.zeros(rps, dtype=np.int32))
    import ctypes as C

    def vp(a):
        return a.ctypes.data_as(C.c_void_p)

    def run_host(i):
        (_, p), (_, o), (_, l) = e2e_bufs[i]
        rc = d.lib.dnab_viterbi_batch(dec._h, rps, vp(p), vp(o), vp(l), vp(h_ll[1]), vp(h_dec[1]), stride,
                                      vp(h_declen[1]), vp(h_status[1]), None, 0, None)
        assert rc == 0, d.lib.dnab_last_error()
    run_host(0)  # warm the staging buffers
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        run_host(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_cells = sum(cells_of(t.n_states, t.k, batches[n_batches - e2e_steps + i][3]) for i in range(e2e_steps))
    h2d = int(np.mean([b[0][1].nbytes + b[1][1].nbytes + b[2][1].nbytes for b in e2e_bufs]))
    d2h = int(h_ll[1].nbytes + h_dec[1].nbytes + h_declen[1].nbytes + h_status[1].nbytes)
    gpu_e2e_decoded = [bytes(h_dec[1][r * stride:r * stride + h_declen[1][r]]).decode("latin1") for r in range(rps)]

    # ---- reduce over ranks: max time, summed work -------------------------------------------------------
    vals = torch.tensor([elapsed_ms, e2e_s, st["timed_fill_ms"], st["timed_traceback_ms"]], dtype=torch.float64, device=dev)
    sums = torch.tensor([my_cells, my_reads, e2e_cells, float(gpu_launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_s, fill_ms, tb_ms = vals.tolist()
    tot_cells, tot_reads, tot_e2e_cells, tot_launches = sums.tolist()

    if rank == 0:
        peak, peak_src = measured_peak()
        # roofline of the fill kernel on rank 0 (per launch)
        fl = max(int(st["timed_fill_launches"]), 1)
        algo_bytes = my_cells + sum(int(np.sum((batches[b][3].astype(np.int64) + 3) // 4)) for b in range(args.warmup, n_batches)) \
            + int(dec_len_last.sum()) * args.steps + 8 * my_reads
        achieved = algo_bytes / (st["timed_fill_ms"] * 1e-3) / 1e9
        fill_kernel = ("viterbiFillBatchKernel" if binfo["enabled"] else
                       "viterbiFillKernel" if args.partition >= 10 else "viterbiFillPushKernel")
        # DRAM traffic of one launch: measured once with `ncu --set full` (profiles/), scaled to this launch
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r01_fill_traffic.json")
        if os.path.exists(tp) and args.workload == "cfg2":
            tj = json.load(open(tp))
            if tj.get("kernel") == fill_kernel:
                traffic = tj["dram_over_algorithmic"] * algo_bytes / fl
                traffic_src = tj["capture"]
        # CPU baseline on a bounded sample of the last batch: one core, decoded strings must match the GPU's
        cpu = None
        if args.cpu_sample > 0:
            sample = batches[-1][0][:args.cpu_sample]
            dt, cpu_dec, kind = cpu_reference_run(w, machine, compiled, sample, 1)
            assert cpu_dec == gpu_e2e_decoded[:len(sample)], "GPU decoded strings differ from the CPU reference"
            cpu_cells = cells_of(t.n_states, t.k, np.array([len(r) for r in sample]))
            cpu = dict(value=cpu_cells / dt, unit=UNIT, cores=1, kind=kind, reads_per_sec=len(sample) / dt,
                       sample=f"first {len(sample)} reads of the last timed batch, single thread, "
                              f"decoded strings identical to the GPU's; host has {os.cpu_count()} cores")
        line = dict(
            metric=METRIC, value=tot_cells / (elapsed_ms * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=elapsed_ms / args.steps, higher_is_better=True, scaling="weak",
            vs_baseline=None, dtype="f64", data="synthetic",
            reads_per_sec=tot_reads / (elapsed_ms * 1e-3),
            config=dict(workload=f"{args.workload}: {w['desc']}, --error-global, -l {w['length']}",
                        reads_per_step_per_gpu=rps, n_states=int(t.n_states), k=int(t.k),
                        **(dict(kernel="read-batched (viterbi_fill_batch.cu): 32 reads per group are the SIMD lanes",
                                team_size=binfo["team_size"], states_per_cta=binfo["states_per_cta"],
                                threads_per_cta=32 * binfo["warps_per_cta"], smem_bytes_per_cta=binfo["smem_bytes_per_cta"],
                                reads_in_flight=32 * binfo["n_teams"],
                                cross_cta_transition_fraction=round(binfo["cross_cta_transition_fraction"], 4))
                           if binfo["enabled"] else
                           dict(kernel="one read per cluster (viterbi_fill_push.cu)",
                                cluster_size=info["cluster_size"], states_per_cta=info["states_per_cta"],
                                threads_per_cta=info["threads_per_cta"], smem_bytes_per_cta=info["smem_bytes_per_cta"],
                                t_in_smem=info["t_in_smem"], table_in_smem=info["table_in_smem"],
                                reads_in_flight=info["n_clusters"])),
                        l2="working set >> L2: every step streams reads_per_step x ~37 MB of predecessor records "
                           "and uses a distinct read batch" if args.workload == "cfg2" else
                           "distinct read batch per step; predecessor-record stream exceeds L2",
                        all_status_ok=status_ok),
            e2e=dict(value=tot_e2e_cells / e2e_s, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                     steps=e2e_steps, reads_per_sec=rps * e2e_steps * world / e2e_s),
            gpu_launches=int(tot_launches),
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                          traffic=traffic, traffic_source=traffic_src,
                          kernel=fill_kernel, launches=fl, avg_launch_ms=st["timed_fill_ms"] / fl,
                          peak_source=peak_src, algorithmic_bytes_per_launch=algo_bytes / fl,
                          fill_share_of_step=fill_ms / elapsed_ms, traceback_share_of_step=tb_ms / elapsed_ms),
            cpu_baseline=cpu, clocks=clocks)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
