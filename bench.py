#!/usr/bin/env python
"""bench.py -- genome sites called per second on N B200s (BASELINE.json metric), next to the host-CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload at every N (config.workload): BASELINE.json configs[1] -- the likelihood-kernel microbench: 1e9 synthetic
per-site count vectors PER GPU fed straight to the genotype / methylation model (weak scaling: rank r owns sites
[r*1e9, (r+1)*1e9) of the same counter-based stream, no data-path collective).  A step = one pass of the hot path
over the rank's 1e9 resident sites (16 launches of k_call_sites over 62.5 M-site slabs; outputs go to one slab-sized
ring in HBM because 1e9 x 200 B does not fit beside the 105 GB of input).

  value     sites/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e       same metric through the C ABI with HOST (pinned) buffers: H2D + kernel + D2H inside the timed region
  roofline  k_call_sites: 306 algorithmic B/site (104 pileup + 1 ref + 200 gt_meth + 1 skip) / event time, against the
            measured HBM copy bandwidth in MEASURED_PEAKS.json
  block_path  segments -> pileup -> model on a synthetic 30x WGBS window (config 3 shape): the default two-kernel
            path, the pileup kernel alone, and the single fused kernel, each with its own roofline
  cpu_baseline / --impl reference: the reference's own calc_gt_prob()/fisher() objects (oracle/_ref, else the oracle
            port) on all host cores over a bounded sample of the same site stream
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018
MEAN_DEPTH = 30.0
BYTES_PER_SITE = 104 + 1 + 200 + 1


_T0 = time.time()


BLOCK_OVERLAP_DEFAULT = False          # what libbsgpu does without BSGPU_BLOCK_OVERLAP in the environment


def log(msg):
    """progress on stderr (the JSON line is the only thing on stdout)"""
    if os.environ.get("RANK", "0") == "0":
        sys.stderr.write("[bench %7.1fs] %s\n" % (time.time() - _T0, msg))
        sys.stderr.flush()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_path(n_sample, nthreads):
    """The CPU implementation of the path over sites [0, n_sample) of the stream; returns (seconds, kind, sample)."""
    from oracle.bindings import Oracle, Reference, reference_available
    o = Oracle()
    p, r = o.synth_sites(SEED, 0, n_sample, MEAN_DEPTH, nthreads=nthreads)
    if reference_available():
        impl, kind = Reference(), "reference"
        what = "reference calc_gt_prob()/fisher() objects (oracle/_ref) + summarise glue"
    else:
        impl, kind = o, "port"
        what = "oracle port (oracle/bs_oracle.c)"
    t0 = time.perf_counter()
    out, skip = impl.call_sites(p, r, nthreads=nthreads)
    dt = time.perf_counter() - t0
    called = int((skip == 0).sum())
    return dt, kind, "%d sites (%d called) of the same counter-based stream, %s, %d threads strided like the reference's calc threads" % (
        n_sample, called, what, nthreads), called, (p, r, out, skip)


def calibrate_cpu_sample(nthreads, target_s):
    dt, _, _, _, _ = cpu_path(200000, nthreads)
    rate = 200000 / max(dt, 1e-6)
    n = int(min(max(rate * target_s, 1e6), 3.2e7))        # <= 10 GB of host records
    return n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nthreads = os.cpu_count() or 1
    n = calibrate_cpu_sample(nthreads, 4.0)
    for _ in range(args.warmup):
        cpu_path(min(n, 2000000), nthreads)
    tot_t, tot_n, kind, sample = 0.0, 0, None, None
    for _ in range(args.steps):
        dt, kind, sample, called, _ = cpu_path(n, nthreads)
        tot_t += dt
        tot_n += called
    value = tot_n / tot_t
    line = {"impl": "reference", "metric": "genome sites called/sec", "value": value, "unit": "sites/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[1]: likelihood microbench, synthetic per-site count vectors; CPU arm runs a bounded sample per step",
                       "sites_per_step": n, "mean_depth": MEAN_DEPTH, "seed": SEED},
            "cpu_baseline": {"value": value, "unit": "sites/s", "cores": nthreads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa_node(torch, local):
    """Run this rank (and so the pinned buffers it allocates and the reader's host threads it starts) on the CPUs of the
    NUMA node its GPU hangs off.  Best effort: returns a note for the JSON line."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "gpu %s reports no numa node" % bus
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "no allowed cpu on numa node %d" % node
        os.sched_setaffinity(0, cpus)
        return "bound to numa node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:
        return "not bound (%s)" % type(e).__name__


def bam_path(args, gpu, bslib, torch, np, stream, rank, world, local):
    """Synthetic coordinate-sorted paired-end WGBS record stream (30x, 150 bp, 5 % positional duplicates, coverage gaps
    every ~100 kb so that the reader cuts blocks) through bsgpu_call_bam with pinned host buffers: H2D of the records,
    decode, block building, normalisation, pileup, model and D2H of gt_vcf[] all inside the timed region."""
    from bs_call_b200.records import GT_VCF
    L, depth, frag, tpb, gap = 150, 30.0, 300, 10000, 400
    sz = int(args.bam_sites)
    nt = int(sz * depth / (2 * L))
    step = 2 * L / depth
    dev = torch.device("cuda", local)
    t = torch.arange(nt, device=dev, dtype=torch.int64)
    p = 101 + torch.floor(t.double() * step).long() + (t // tpb) * gap
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + 17 + rank)
    dup = (torch.rand(nt, device=dev, generator=g) < 0.05) & (t > 0)
    src = torch.where(dup, t - 1, t)
    dlen = (frag - L) + ((src * 2654435761) >> 7) % 41 - 20
    pos_f = p[src]
    pos_r = pos_f + dlen
    order = torch.argsort(torch.cat([pos_f, pos_r]), stable=True)
    rank_of = torch.empty_like(order)
    rank_of[order] = torch.arange(2 * nt, device=dev)
    ctg_len = int(pos_r.max().item()) + L + 600
    nbytes = gpu.synth_bam_bytes(nt, L)
    d_bam = torch.empty(nbytes + 16, dtype=torch.uint8, device=dev)
    d_ref = torch.empty(ctg_len + 16, dtype=torch.uint8, device=dev)
    u32 = lambda a: a.to(torch.int32).contiguous()
    a_f, a_r, a_s, a_k = u32(pos_f), u32(pos_r), u32(src), u32(rank_of)
    gpu.synth_bam_dev(SEED, nt, L, a_f.data_ptr(), a_r.data_ptr(), a_s.data_ptr(), a_k.data_ptr(), d_bam.data_ptr(), stream)
    gpu.synth_ref_dev(SEED, 1, ctg_len, d_ref.data_ptr(), stream)
    torch.cuda.synchronize()
    hbam = bslib.HostBuffer(nbytes, np.uint8)
    hvcf = bslib.HostBuffer(ctg_len + 8, GT_VCF)
    hbam.array[:] = d_bam[:nbytes].cpu().numpy()
    href = d_ref[:ctg_len].cpu().numpy()
    del d_bam, d_ref
    tl = np.array([ctg_len], dtype=np.uint32)
    for _ in range(2):
        blocks, vcf = gpu.call_bam(hbam.array, tl, [href], vcf=hvcf.array)
    s0 = gpu.stats()
    steps = 3                                       # calls timed per figure of this leg, whatever K is (a call is 25-50 ms)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        blocks, vcf = gpu.call_bam(hbam.array, tl, [href], vcf=hvcf.array)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    s1 = gpu.stats()
    called = int((vcf["skip"] == 0).sum())
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    cc = torch.tensor([called], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
    out = {"workload": "synthetic %dx paired-end %d-bp WGBS record stream over %d sites (config 3 shape): %d records, %d blocks, 5 %% duplicates" % (
               int(depth), L, sz, 2 * nt, len(blocks)),
           "e2e": {"value": float(cc.item()) / float(tt.item()), "unit": "sites/s", "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) // steps,
                   "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"]) // steps, "sites_called_per_step": called,
                   "records_per_s": 2 * nt / dt, "note": "bsgpu_call_bam on pinned host buffers, per rank work fixed"},
           "stage_s_per_step": {"frame_h2d_decode": (s1["bam_decode_s"] - s0["bam_decode_s"]) / steps,
                                "descriptors_d2h_block_builder_host": (s1["bam_build_s"] - s0["bam_build_s"]) / steps,
                                "normalise_pileup_model_d2h": (s1["bam_call_s"] - s0["bam_call_s"]) / steps},
           "gpu_launches_per_step": (s1["kernel_launches"] - s0["kernel_launches"]) // steps}
    log("bam: gt_vcf %.3g sites/s; to BCF ..." % out["e2e"]["value"])
    # the same stream all the way to BCF records (the writer's derivations on the device): only the records come home
    hbcf = bslib.HostBuffer(ctg_len * 96 + 4096, np.uint8)
    for _ in range(2):
        _, rb, rn = gpu.call_bam_bcf(hbam.array, tl, [href], out=hbcf.array)
    sb0 = gpu.stats()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        _, rb, rn = gpu.call_bam_bcf(hbam.array, tl, [href], out=hbcf.array)
    torch.cuda.synchronize()
    dtb = (time.perf_counter() - t0) / steps
    sb1 = gpu.stats()
    tb_ = torch.tensor([dtb], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tb_, op=dist.ReduceOp.MAX)
    out["to_bcf_records"] = {"value": float(cc.item()) / float(tb_.item()), "unit": "sites/s", "records_per_step": rn, "record_bytes_per_step": len(rb),
                             "h2d_bytes_per_step": (sb1["h2d_bytes"] - sb0["h2d_bytes"]) // steps, "d2h_bytes_per_step": (sb1["d2h_bytes"] - sb0["d2h_bytes"]) // steps,
                             "stage_s_per_step": {"frame_h2d_decode": (sb1["bam_decode_s"] - sb0["bam_decode_s"]) / steps,
                                                  "descriptors_d2h_block_builder_host": (sb1["bam_build_s"] - sb0["bam_build_s"]) / steps,
                                                  "normalise_pileup_model_writer_d2h": (sb1["bam_call_s"] - sb0["bam_call_s"]) / steps},
                             "gpu_launches_per_step": (sb1["kernel_launches"] - sb0["kernel_launches"]) // steps,
                             "note": "bsgpu_call_bam_bcf on pinned host buffers: BAM records up, BCF records down"}
    bcf_first = None
    if rank == 0 and world == 1 and not args.no_cpu:
        bcf_first = rb.copy()
    hbcf.free()
    # the same with the --report-file side channels on (conversion profile + base / read tallies by the normalisation
    # kernel, read_input's tallies by the host builder)
    gpu.profile_enable(True)
    gpu.call_bam(hbam.array, tl, [href], vcf=hvcf.array)
    gpu.profile_read(reset=True)
    sp0 = gpu.stats()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        gpu.call_bam(hbam.array, tl, [href], vcf=hvcf.array)
    torch.cuda.synchronize()
    dtp = (time.perf_counter() - t0) / steps
    sp1 = gpu.stats()
    prof = gpu.profile_read(reset=True)
    tp_ = torch.tensor([dtp], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tp_, op=dist.ReduceOp.MAX)
    out["with_report_side_channels"] = {"value": float(cc.item()) / float(tp_.item()), "unit": "sites/s", "slowdown": dtp / dt,
                                         "normalise_pileup_model_d2h_s": (sp1["bam_call_s"] - sp0["bam_call_s"]) / steps,
                                         "gpu_launches_per_step": (sp1["kernel_launches"] - sp0["kernel_launches"]) // steps,
                                         "profile_counts_per_step": int(prof["conv"].sum()) // steps, "profile_used": prof["used"]}
    # The reference's own chain (read_input -> process_template_vector -> call_genotypes_ML, all its threads) and its writer:
    # timed on a bounded prefix of the stream (cpu_baseline), then run over --cpu-diff-records of it (default: ALL, config 3
    # "VCF/BCF diffed against the CPU run") with its stats live, and every block's gt_vcf[] records, every BCF record and the
    # --report-file side channels compared with what the device produced for the same records.
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.bindings import Oracle, Reference, bcf_diff, reference_available
        from tests import blockgen, util
        rb = nbytes // (2 * nt)
        ncores = os.cpu_count() or 1
        if reference_available():
            impl, kind = Reference(calc_threads=ncores), "reference"
            what = "reference read_input -> process_template_vector -> call_genotypes_ML (oracle/_ref/libbsref.so)"
        else:
            impl, kind = Oracle(), "port"
            what = "oracle port of the same chain (1 thread)"

        def prefix_of(nrec_c):
            pre = hbam.array[:nrec_c * rb]
            last_pos = int(pre[(nrec_c - 1) * rb + 8:(nrec_c - 1) * rb + 12].view("<i4")[0]) + 1
            return pre, min(ctg_len, last_pos + 2 * L + 1200)

        log("bam: BCF %.3g sites/s; CPU reference on a prefix, then the diff ..." % out["to_bcf_records"]["value"])
        nrec_t = int(min(2 * nt, 400000))
        pre, clen = prefix_of(nrec_t)
        t0 = time.perf_counter()
        cb, ct, _, _, cv = impl.read_input(pre, [clen], [href[:clen]], run_chain=True)
        cs = time.perf_counter() - t0
        ccalled = int((cv["skip"] == 0).sum())
        out["cpu_baseline"] = {"value": ccalled / cs, "unit": "sites/s", "cores": ncores, "kind": kind,
                               "sample": "first %d records (%d sites called, %d blocks) of the same stream; %s" % (nrec_t, ccalled, len(cb), what)}
        del cb, cv
        nrec_c = int(min(2 * nt, args.cpu_diff_records if args.cpu_diff_records > 0 else 2 * nt))
        whole = nrec_c == 2 * nt
        pre, clen = prefix_of(nrec_c)
        if kind == "reference":
            impl.stats_enable(True); impl.stats_reset()
        else:
            impl.profile_enable(True); impl.profile_reset()
        t0 = time.perf_counter()
        cb, ct, _, _, cv = impl.read_input(pre, [clen], [href[:clen]], run_chain=True)
        cs_all = time.perf_counter() - t0
        if kind == "reference":
            cprof = impl.stats_read(); impl.stats_enable(False)
        else:
            cprof = impl.profile_read(); impl.profile_enable(False)
        log("bam: CPU chain over %d records took %.1f s; comparing ..." % (nrec_c, cs_all))
        # every block of the whole stream (every complete block of a prefix: its last block is cut short)
        nblk = len(cb) if whole else len(cb) - 1
        assert whole is False or len(cb) == len(blocks)
        checked = 0
        for b, w in zip(blocks[:nblk], cb[:nblk]):
            assert (b["x"], b["y"], b["n_templates"]) == (w["x"], w["y"], w["n_templates"])
            n_ = int(w["y"]) - int(w["x"]) + 1
            checked += util.assert_vcf_close(vcf[int(b["vcf_off"]):int(b["vcf_off"]) + n_], cv[int(w["vcf_off"]):int(w["vcf_off"]) + n_])
        out["cpu_diff"] = {"records_in": nrec_c, "whole_stream": whole, "blocks": nblk, "sites_called_checked": checked,
                           "cpu_chain_s": cs_all, "against": what + ", stats live"}
        # the records: the reference's writer over the reference's own chain, block by block, against the device's stream
        if kind == "reference":
            want = []
            t0 = time.perf_counter()
            for w in cb[:nblk]:
                n_ = int(w["y"]) - int(w["x"]) + 1
                wb_, _ = impl.print_block(cv[int(w["vcf_off"]):int(w["vcf_off"]) + n_], blockgen.window_codes(href[:clen], int(w["x"]), int(w["y"]) + 2),
                                          int(w["x"]), rid=0, ctg_end=ctg_len)
                want.append(wb_)
            ws = time.perf_counter() - t0
            want = np.concatenate(want)
            got = bcf_first if whole else bcf_first[:len(want)]
            d = bcf_diff(got, want)
            assert d["records_a"] == d["records_b"] == d["fixed_equal"] and d["order_violations"] == 0, d
            out["cpu_diff"].update({"bcf_records": d["records_a"], "bcf_records_fixed_fields_equal": d["fixed_equal"],
                                    "bcf_records_byte_identical": d["identical"], "bcf_bytes": len(want), "cpu_writer_s": ws})
            out["to_bcf_records"]["parity_records_checked"] = d["records_a"]
            out["to_bcf_records"]["parity_records_byte_identical"] = d["identical"]
            del want
        del cv
        log("bam: records compared; side channels ...")
        gpu.call_bam(pre, [clen], [href[:clen]], vcf=hvcf.array)
        util.same_profile(gpu.profile_read(reset=True), cprof, "bench stream", recycled_vectors=kind == "reference")
        out["with_report_side_channels"]["parity_profile_counts_checked"] = int(cprof["conv"].sum())
    gpu.profile_enable(False)
    hbam.free()
    hvcf.free()
    return out


def genome_path(args, gpu, bslib, torch, np, stream, rank, world, local):
    """BASELINE.json configs[4] shape: a 24-contig hg38-shaped genome (contig lengths / --genome-scale), 30x paired-end
    WGBS record streams, sharded over the ranks by shard.plan (level 1: contigs by longest-processing-time; level 2: the
    heaviest contigs cut in two AT BLOCK BOUNDARIES) -- strong scaling, the genome is the same at every N.  Every rank
    pushes its regions through ONE streaming session (bsgpu_bam_open / _feed / _cut / _drain): BAM records in pinned host
    memory in, BCF records in pinned host memory out, H2D / D2H inside the timed region, a feeding thread and a printing
    thread as in the reference.  Then the ordered merge: the per-batch extents of all ranks gathered on rank 0 and put in
    coordinate order (the reference's workflow is one output per region, concatenated: src/process_sam_header.c:52-70)."""
    import threading
    import zlib
    from bs_call_b200 import synthgenome as sg
    import torch.distributed as dist
    dev = torch.device("cuda", local)
    scale = int(args.genome_scale)
    rb = gpu.synth_bam_bytes(1, sg.READ_LEN) // 2          # bytes of one record (a template is two)
    if scale <= 0:
        # largest genome whose records (in) and BCF records (out) for ONE rank holding everything fit in a quarter of the host's free memory
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 64 << 30
        scale = 64
        for sc in (8, 16, 32):
            nt_tot = sum(sg.n_templates(ln) for ln in sg.contig_lengths(sc))
            if nt_tot * 2 * rb * 2.2 < avail / 4:
                scale = sc
                break
    lens, plan = sg.plan(scale, world)
    mine = [(r,) + sg.region_templates(r, lens[r.contig]) for r in plan[rank]]
    mine = [(r, t0, t1) for r, t0, t1 in mine if t1 > t0]
    nbytes_in = sum((t1 - t0) * 2 * rb for _, t0, t1 in mine)
    hbam = bslib.HostBuffer(max(nbytes_in, 16), np.uint8)
    contigs = sorted({r.contig for r, _, _ in mine})
    href = {c: bslib.HostBuffer(lens[c], np.uint8) for c in contigs}
    d_ref = torch.empty(max(lens) + 16, dtype=torch.uint8, device=dev)
    for c in contigs:
        gpu.synth_ref_dev(SEED + 1000003 * c, 1, lens[c], d_ref.data_ptr(), stream)
        torch.cuda.synchronize()
        torch.from_numpy(href[c].array).copy_(d_ref[:lens[c]])
    del d_ref
    spans, at = [], 0
    max_reg = max([(t1 - t0) * 2 * rb for _, t0, t1 in mine] + [16])
    d_out = torch.empty(max_reg + 16, dtype=torch.uint8, device=dev)
    for r, t0, t1 in mine:
        n = sg.region_stream_dev(gpu, torch, dev, SEED, r.contig, t0, t1, d_out, stream)
        view = hbam.array[at:at + n]
        torch.from_numpy(view).copy_(d_out[:n])
        sg.patch_contig(view, r.contig, rb)
        spans.append((r, at, at + n))
        at += n
    del d_out
    torch.cuda.empty_cache()
    log("genome: %d regions, %.2f GB of records on this rank, scale 1/%d" % (len(spans), nbytes_in / 1e9, scale))
    codes = [href[c].array if c in href else None for c in range(len(lens))]
    # Lanes: a session runs its batches one after the other (frame -> upload -> decode -> blocks -> windows), so ONE session
    # leaves the device idle during the host stages of a batch.  --genome-sessions K opens K contexts + sessions on the GPU
    # and deals the rank's regions out to them (longest first): the host stages of one lane run under the device stages of
    # another.  Results stay per region, so the ordered merge is the same.
    from bs_call_b200 import shard as _shard
    nl = max(1, min(int(args.genome_sessions), len(spans)))
    owner = _shard.lpt_assign([hi - lo for _, lo, hi in spans], nl)
    # dbSNP annotation of the rank's contigs (configs[4] is specified with it): ids, and known sites that are written whatever was called
    db = {c: sg.contig_dbsnp(SEED, c, lens[c]) for c in contigs}
    db_entries = sum(len(v[0]) for v in db.values())
    lanes = []
    for k in range(nl):
        g = gpu if k == 0 else bslib.BsGpu(device=local)
        for c in contigs:
            g.set_contig_annotation(c, dbsnp=bslib.dbsnp(*db[c]))
        lanes.append({"gpu": g, "spans": [sp for sp, o in zip(spans, owner) if o == k],
                      "sess": g.bam_session(np.array(lens, dtype=np.uint32), codes, bcf=True, batch_bytes=int(args.genome_batch_mb) << 20)})
    SLICE = 32 << 20

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def lane_pass(lane, res, err):
        sess = lane["sess"]

        def printer():
            try:
                while True:
                    got = sess.drain(wait=True)
                    if got is not None:
                        b, d, n, r = got
                        res.append((b.copy(), d, len(d), n, (sess, r)))
                    elif not sess.finished:
                        time.sleep(0.0002)
                    if sess.finished:
                        return
            except Exception as e:      # noqa: BLE001
                err.append(e)

        th = threading.Thread(target=printer)
        th.start()
        try:
            for r, lo, hi in lane["spans"]:
                for off in range(lo, hi, SLICE):
                    sess.feed(hbam.array[off:min(off + SLICE, hi)])
                sess.cut()                       # a region ends where a block ends: no batch of results mixes two regions
            sess.finish()
        except Exception as e:          # noqa: BLE001
            err.append(e)
        finally:
            th.join()

    def one_pass():
        res, err = [[] for _ in lanes], []
        barrier()
        t0 = time.perf_counter()
        ths = [threading.Thread(target=lane_pass, args=(ln, res[k], err)) for k, ln in enumerate(lanes)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        dt = time.perf_counter() - t0
        if err:
            raise err[0]
        return dt, [x for lst in res for x in lst]

    def give_back(res):
        for _, _, _, _, (sess, r) in res:
            sess.release(r)
        for ln in lanes:
            ln["sess"].rewind()

    def all_stats():
        st = [ln["gpu"].stats() for ln in lanes]
        return {k: sum(x[k] for x in st) for k in st[0]}

    # every pass keeps all its results lent out until it is over (they are the output of the run), so the sessions' pools of
    # pinned result buffers reach their final size during the warm-up passes
    for _ in range(2):
        dtw, res = one_pass()
        log("genome: warm-up pass %.3f s, %d result batches" % (dtw, len(res)))
        give_back(res)
    s0 = all_stats()
    steps = 3
    dts = []
    for k in range(steps):
        dt, res = one_pass()
        dts.append(dt)
        if k < steps - 1:
            give_back(res)
    s1 = all_stats()
    progs = [ln["sess"].progress() for ln in lanes]
    prog = {k: sum(x[k] for x in progs) for k in progs[0]}
    log("genome: timed passes %s s" % ", ".join("%.3f" % v for v in dts))
    dt = sum(dts) / steps
    called = (s1["sites_called"] - s0["sites_called"]) // steps
    # ---- ordered merge (timed): every batch of results is an extent (contig, first block x, last block y, bytes, records)
    barrier()
    tm0 = time.perf_counter()
    table = [(int(b[0]["tid"]), int(b[0]["x"]), int(b[-1]["y"]), nb, nr, rank, i) for i, (b, _, nb, nr, _) in enumerate(res) if len(b)]
    if world > 1:
        tables = [None] * world
        dist.all_gather_object(tables, table)
    else:
        tables = [table]
    merged = sorted(e for t in tables for e in t)
    for a, b2 in zip(merged, merged[1:]):
        assert a[0] < b2[0] or a[2] <= b2[1], "extents overlap: %r %r" % (a, b2)
    merge_s = time.perf_counter() - tm0
    # ---- what must be the same at every N: records and bytes of the whole genome, and the records of the two smallest contigs
    small = sorted(range(len(lens)), key=lambda c: lens[c])[:2]
    crc = {c: 0 for c in small}
    keep_small = {}
    for b, d, nb, nr, r in res:
        if len(b) and int(b[0]["tid"]) in crc:
            c = int(b[0]["tid"])
            crc[c] = zlib.crc32(d, crc[c])
            if rank == 0 and world == 1 and c == small[0]:
                keep_small.setdefault(c, []).append(d.copy())
    tot = torch.tensor([sum(nb for _, _, nb, _, _ in res), sum(nr for _, _, _, nr, _ in res), called, nbytes_in, len(res)] +
                       [crc[c] for c in small] +
                       [(s1[k] - s0[k]) // steps for k in ("h2d_bytes", "d2h_bytes", "kernel_launches")], dtype=torch.int64, device=dev)
    tmax = torch.tensor([dt, max(dts), min(dts)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)       # a contig's records all sit on one rank unless it was split: crc of a split contig is not comparable
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    split_contigs = sorted({r.contig for lst in plan for r in lst if not (r.start == 1 and r.stop == lens[r.contig])})
    loads = [sum(sg.region_templates(r, lens[r.contig])[1] - sg.region_templates(r, lens[r.contig])[0] for r in lst) for lst in plan]
    out = {"workload": "synthetic hg38-shaped genome, 24 contigs at 1/%d scale (%d positions), 30x paired-end %d-bp WGBS records with 5 %% duplicates and a "
                       "coverage gap every ~100 kb, dbSNP annotation (one known site per %d positions, one in %d of them always written); BASELINE.json configs[4] shape" % (scale, sum(lens), sg.READ_LEN, sg.DBSNP_EVERY, sg.DBSNP_ALWAYS),
           "scaling": "strong", "value": float(tot[2].item()) / float(tmax[0].item()), "unit": "sites/s",
           "pass_s": float(tmax[0].item()), "pass_s_minmax": [float(tmax[2].item()), float(tmax[1].item())], "passes_timed": steps,
           "merge_s": merge_s, "value_with_merge": float(tot[2].item()) / (float(tmax[0].item()) + merge_s),
           "sites_called": int(tot[2].item()), "records_in_bytes": int(tot[3].item()), "bcf_records": int(tot[1].item()), "bcf_bytes": int(tot[0].item()),
           "result_batches": int(tot[4].item()), "extents_merged": len(merged),
           "dbsnp_entries_rank0": db_entries,
           "plan": {"regions": sum(len(lst) for lst in plan), "contigs_split_at_block_boundaries": split_contigs,
                    "templates_per_rank": loads, "imbalance": max(loads) / (sum(loads) / len(loads))},
           "invariants": {"crc32_contig_%d" % c: (int(tot[5 + i].item()) if c not in split_contigs else None) for i, c in enumerate(small)},
           "session_rank0": {"sessions_per_gpu": nl, "batch_bytes": int(args.genome_batch_mb) << 20, "batches_per_pass": prog["batches"] // (steps + 2),
                             "carry_bytes_per_pass": prog["carry_bytes"] // (steps + 2), "empty_batches": prog["empty_batches"],
                             "pinned_bytes": prog["pinned_bytes"]},
           "h2d_bytes_per_pass": int(tot[5 + len(small)].item()), "d2h_bytes_per_pass": int(tot[6 + len(small)].item()),
           "gpu_launches_per_pass": int(tot[7 + len(small)].item()),
           "note": "per rank: %d session(s), each fed by bsgpu_bam_feed in 32 MiB slices from pinned host memory + bsgpu_bam_cut per region and drained by a thread of its own; " % nl +
                   "value = sites called on all ranks / max over ranks of the wall time of a pass"}
    # ---- parity: the smallest contig against the reference's own chain + writer, record for record (rank 0 of a 1-GPU run)
    if rank == 0 and world == 1 and not args.no_cpu and keep_small:
        from oracle.bindings import Reference, bcf_diff, reference_available
        from tests import blockgen
        c = small[0]
        r, lo, hi = [sp for sp in spans if sp[0].contig == c][0]
        if reference_available():
            ncores = os.cpu_count() or 1
            impl = Reference(calc_threads=ncores)
            sub = hbam.array[lo:hi].copy()
            sg.patch_contig(sub, 0, rb)
            hc = href[c].array
            t0 = time.perf_counter()
            cb, ct, _, _, cv = impl.read_input(sub, [lens[c]], [hc], run_chain=True)
            cs = time.perf_counter() - t0
            want = []
            t0 = time.perf_counter()
            for w in cb:
                n_ = int(w["y"]) - int(w["x"]) + 1
                wb_, _ = impl.print_block(cv[int(w["vcf_off"]):int(w["vcf_off"]) + n_], blockgen.window_codes(hc, int(w["x"]), int(w["y"]) + 2),
                                          int(w["x"]), rid=c, ctg_end=lens[c], dbsnp=db[c])
                want.append(wb_)
            ws = time.perf_counter() - t0
            diff = bcf_diff(np.concatenate(keep_small[c]), np.concatenate(want))
            assert diff["records_a"] == diff["records_b"] == diff["fixed_equal"] and diff["order_violations"] == 0, diff
            ccalled = int((cv["skip"] == 0).sum())
            out["parity"] = {"contig": c, "positions": lens[c], "sites_called": ccalled, "blocks": len(cb), "records": diff["records_a"],
                             "records_fixed_fields_equal": diff["fixed_equal"], "records_byte_identical": diff["identical"],
                             "against": "reference read_input -> process_template_vector -> call_genotypes_ML -> print_vcf_entry (oracle/_ref/libbsref.so)"}
            out["cpu_baseline"] = {"value": ccalled / (cs + ws), "unit": "sites/s", "cores": ncores, "kind": "reference",
                                   "sample": "contig %d of the same genome (%d positions, %d sites called): the reference's chain on all cores (%.2f s) + its writer on one thread (%.2f s)" % (c, lens[c], ccalled, cs, ws)}
    for _, _, _, _, (sess, r) in res:
        sess.release(r)
    for k, ln in enumerate(lanes):
        ln["sess"].close()
        if k:
            ln["gpu"].close()
    hbam.free()
    for h in href.values():
        h.free()
    return out


def full_binary_path(args, gpu, bslib, torch, np, stream, local):
    """The whole program: the reference's unmodified bs_call binary and the same program with the product's
    seam files (oracle/_ref/bs_call, bs_call_gpu: oracle/Makefile, over oracle/minihts as htslib stand-in), both run from the
    command line on a BAM file + FASTA file of the config-1 shape (BASELINE.json configs[0] and [2]: "bs_call -L5", the same BAM
    end to end on one B200, BCF diffed against the CPU run); wall clock of the processes, options as SURVEY.md 8d states them.
    rank 0, N = 1 only."""
    import subprocess, tempfile, shutil
    from bs_call_b200 import hostio
    from oracle.bindings import bcf_diff
    here = os.path.dirname(os.path.abspath(__file__))
    refdir = os.path.join(here, "oracle", "_ref")
    bins = {k: os.path.join(refdir, k) for k in ("bs_call", "bs_call_gpu")}
    for b in bins.values():
        if not os.path.exists(b):
            raise RuntimeError("%s not built (the reference tree was absent when oracle/Makefile ran)" % b)
    L, depth, frag, tpb, gap = 150, 30.0, 300, 10000, 400
    sz = int(args.binary_sites)
    nt = int(sz * depth / (2 * L))
    step = 2 * L / depth
    dev = torch.device("cuda", local)
    t = torch.arange(nt, device=dev, dtype=torch.int64)
    p = 101 + torch.floor(t.double() * step).long() + (t // tpb) * gap
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + 29)
    dup = (torch.rand(nt, device=dev, generator=g) < 0.05) & (t > 0)
    src = torch.where(dup, t - 1, t)
    dlen = (frag - L) + ((src * 2654435761) >> 7) % 41 - 20
    pos_f = p[src]
    pos_r = pos_f + dlen
    order = torch.argsort(torch.cat([pos_f, pos_r]), stable=True)
    rank_of = torch.empty_like(order)
    rank_of[order] = torch.arange(2 * nt, device=dev)
    ctg_len = int(pos_r.max().item()) + L + 600
    nbytes = gpu.synth_bam_bytes(nt, L)
    d_bam = torch.empty(nbytes + 16, dtype=torch.uint8, device=dev)
    d_ref = torch.empty(ctg_len + 16, dtype=torch.uint8, device=dev)
    u32 = lambda a: a.to(torch.int32).contiguous()
    a_f, a_r, a_s, a_k = u32(pos_f), u32(pos_r), u32(src), u32(rank_of)
    gpu.synth_bam_dev(SEED, nt, L, a_f.data_ptr(), a_r.data_ptr(), a_s.data_ptr(), a_k.data_ptr(), d_bam.data_ptr(), stream)
    gpu.synth_ref_dev(SEED, 1, ctg_len, d_ref.data_ptr(), stream)
    torch.cuda.synchronize()
    hbam = d_bam[:nbytes].cpu().numpy()
    href = d_ref[:ctg_len].cpu().numpy()
    del d_bam, d_ref
    tl = np.array([ctg_len], dtype=np.uint32)
    # sites called on this stream (pileup.n > 0): the library's own counter over one in-memory pass
    hbcf = bslib.HostBuffer(ctg_len * 96 + 4096, np.uint8)
    s0 = gpu.stats()
    gpu.call_bam_bcf(hbam, tl, [href], out=hbcf.array)
    called = gpu.stats()["sites_called"] - s0["sites_called"]
    hbcf.free()                                     # page-locked: give it back before the child processes lock their own
    tmp = tempfile.mkdtemp(prefix="bsgpu_fullbin_", dir=os.environ.get("BENCH_TMPDIR"))
    try:
        fa, bf = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "in.bam")
        t0 = time.perf_counter()
        hostio.write_fasta(fa, ["ctg0"], [href])
        hostio.write_bam(bf, ["ctg0"], tl, hbam, level=0)
        prep = time.perf_counter() - t0
        ncores = os.cpu_count() or 1

        def run(tag, binary, env_extra):
            outp = os.path.join(tmp, tag + ".bcf")
            cmd = [binary, "-r", fa, "-n", "S", "-L", "5", "--benchmark-mode", "-O", "u", "-o", outp, "-t", "%d,0,0" % max(1, ncores - 1), bf]
            env = dict(os.environ)
            env.update(env_extra)
            env["BSGPU_DEVICE"] = str(local)
            t0 = time.perf_counter()
            pr = subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
            err = pr.stderr.read().decode(errors="replace")
            _, status, ru = os.wait4(pr.pid, 0)
            wall = time.perf_counter() - t0
            pr.returncode = os.waitstatus_to_exitcode(status)
            if pr.returncode != 0:
                raise RuntimeError("%s failed (%d): %s" % (tag, pr.returncode, err[-600:]))
            thr = [ln for ln in err.split("\n") if ln.startswith("Additional threads:")]
            stamps = [ln[len("bsgpu seam:"):].strip() for ln in err.split("\n") if ln.startswith("bsgpu seam:")]      # BSGPU_SEAM_TIMING=1
            return {**({"seam_timing": stamps} if stamps else {}),"value": called / wall, "unit": "sites/s", "wall_s": wall, "user_s": ru.ru_utime, "sys_s": ru.ru_stime,
                    "maxrss_mb": ru.ru_maxrss / 1024.0, "additional_threads_calc_input_output": thr[0].split(":")[1].split() if thr else None,
                    "cmd": " ".join(os.path.basename(c) if c.startswith(tmp) or c.startswith(refdir) else c for c in cmd)}, outp

        out = {"workload": "config-1 shape at %d sites: synthetic 30x paired-end 150-bp WGBS, %d records (%.2f GB of BAM records, BGZF level 0), one contig of %d bp, 5 %% duplicates, coverage gap every ~100 kb; FASTA + .fai; output uncompressed BCF" % (
                   sz, 2 * nt, nbytes / 1e9, ctg_len),
               "sites_called": int(called), "host_cores": ncores, "file_prep_s": prep,
               "htslib": "oracle/minihts (this repository's stand-in: no htslib in the image; single-threaded BGZF, BAM, faidx, BCF2), so every additional thread goes to the calc threads (-t n,0,0; src/parse_args.c:191-213 would give 3/7 of them to BGZF input threads the stand-in does not have)"}
        log("full binary: reference bs_call on %d cores ..." % ncores)
        def best_of(n, *a):
            # process start-up (CUDA context, page cache) varies by seconds from run to run: the faster of n runs, all walls reported
            runs = [run(*a) for _ in range(n)]
            best = min(runs, key=lambda r: r[0]["wall_s"])
            best[0]["wall_s_all_runs"] = [r[0]["wall_s"] for r in runs]
            return best

        out["cpu_reference_binary"], f_cpu = best_of(2, "cpu", bins["bs_call"], {})
        log("full binary: bs_call_gpu (seam C, then seam D) ...")
        out["gpu_seam_C"], f_c = best_of(2, "gpu_c", bins["bs_call_gpu"], {})
        out["gpu_seam_D"], f_d = best_of(2, "gpu_d", bins["bs_call_gpu"], {"BSGPU_SEAM_RECORDS": "1"})
        out["gpu_seam_C"]["note"] = "read_input / process_template_vector / call_genotypes_ML on the device, the reference's print thread writes (src/print_vcf.c on one host thread)"
        out["gpu_seam_D"]["note"] = "additionally print_vcf_entry's work on the device: BCF records handed to bcf_write"
        _, rc_ = hostio.read_bcf(f_cpu)
        par = {}
        for tag, f in (("seam_C", f_c), ("seam_D", f_d)):
            _, rg_ = hostio.read_bcf(f)
            d = bcf_diff(rg_, rc_)
            par[tag] = {"records": d["records_a"], "records_cpu": d["records_b"], "fixed_fields_equal": d["fixed_equal"], "byte_identical": d["identical"]}
            assert d["records_a"] == d["records_b"] and d["order_violations"] == 0, d
            assert d["identical"] >= d["records_a"] - max(3, d["records_a"] // 100000), d
        out["parity"] = par
        out["speedup_wall"] = {"seam_C": out["cpu_reference_binary"]["wall_s"] / out["gpu_seam_C"]["wall_s"],
                               "seam_D": out["cpu_reference_binary"]["wall_s"] / out["gpu_seam_D"]["wall_s"]}
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def writer_path(args, gpu, bslib, torch, np, stream, rank, world, local):
    """the writer's per-site derivations on the device.  (a) gt_vcf[] resident -> BCF records
    resident (the three writer kernels alone); (b) count vectors on the host -> BCF records on the host
    (bsgpu_call_sites_bcf: H2D, model, writer, D2H of the records only); (c) the reference's own writer
    (print_vcf_entry / flush_vcf_entries, one thread as in the reference) on a bounded sample, with byte parity."""
    from bs_call_b200.records import GT_VCF, PILEUP
    n = int(args.e2e_sites)
    dev = torch.device("cuda", local)
    d_pile = torch.empty(n * 104 + 16, dtype=torch.uint8, device=dev)
    d_ref = torch.empty(n + 16, dtype=torch.uint8, device=dev)
    d_vcf = torch.empty(n * 208 + 16, dtype=torch.uint8, device=dev)
    gpu.synth_sites_dev(SEED, rank * n, n, MEAN_DEPTH, d_pile.data_ptr(), d_ref.data_ptr(), stream)
    d_ref[n:n + 2] = 1
    gpu.call_sites_vcf_dev(d_pile.data_ptr(), d_ref.data_ptr(), n, d_vcf.data_ptr(), stream)
    torch.cuda.synchronize()
    cap = n * 160
    d_out = torch.empty(cap + 16, dtype=torch.uint8, device=dev)
    for _ in range(2):
        nb, nr = gpu.bcf_block_dev(d_vcf.data_ptr(), d_ref.data_ptr(), 1, n, d_out.data_ptr(), cap, stream=stream)
    steps = max(3, args.steps)
    l0 = gpu.stats()["kernel_launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        nb, nr = gpu.bcf_block_dev(d_vcf.data_ptr(), d_ref.data_ptr(), 1, n, d_out.data_ptr(), cap, stream=stream)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    peak, peak_kind = measured_peaks()
    alg = 208.0 * n + nb                              # every gt_vcf record read once, every record byte written once
    out = {"workload": "%d sites of the config-2 stream as one block: gt_vcf[] -> BCF records (%d records, %.1f B each)" % (n, nr, nb / max(nr, 1)),
           "resident": {"value": n / (ms * 1e-3), "unit": "sites/s", "ms": ms, "gpu_launches": (gpu.stats()["kernel_launches"] - l0) // steps,
                        "roofline": {"kernel": "k_bcf_calls + k_bcf_measure + k_bcf_offsets + k_bcf_emit", "bound": "hbm",
                                     "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                                     "algorithmic_bytes_per_site": alg / n}}}
    # (b) host count vectors -> host records
    hp = bslib.HostBuffer(n, PILEUP)
    hr = bslib.HostBuffer(n + 2, np.uint8)
    ho = bslib.HostBuffer(cap, np.uint8)
    hp.array.view(np.uint8)[:] = d_pile[:n * 104].cpu().numpy()
    hr.array[:] = d_ref[:n + 2].cpu().numpy()
    called = int((hp.array["n"] > 0).sum())
    for _ in range(2):
        rb, rn = gpu.call_sites_bcf(hp.array, hr.array, 1, out=ho.array)
    s0 = gpu.stats()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        rb, rn = gpu.call_sites_bcf(hp.array, hr.array, 1, out=ho.array)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    s1 = gpu.stats()
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    cc = torch.tensor([called], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
    assert rn == nr and len(rb) == nb
    out["e2e"] = {"value": float(cc.item()) / float(tt.item()), "unit": "sites/s", "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) // steps,
                  "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"]) // steps, "records_per_step": rn,
                  "note": "bsgpu_call_sites_bcf on pinned host arrays: count vectors up, BCF records down"}
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.bindings import Oracle, Reference, reference_available
        m = int(min(n, 1500000))
        vcf = d_vcf[:m * 208].cpu().numpy().view(GT_VCF)
        refw = hr.array[:m + 2].copy()
        if reference_available():
            impl, kind, what = Reference(), "reference", "reference print_vcf_entry / flush_vcf_entries / _print_vcf_entry (oracle/_ref/libbsref.so), one thread as in bs_call"
        else:
            impl, kind, what = Oracle(), "port", "oracle port of the writer, one thread"
        t0 = time.perf_counter()
        wb, wn = impl.print_block(vcf, refw, 1)
        cs = time.perf_counter() - t0
        gb, gn = gpu.bcf_block(vcf, refw, 1)
        assert gn == wn and gb.tobytes() == wb.tobytes(), "device writer differs from the %s" % kind
        out["cpu_baseline"] = {"value": m / cs, "unit": "sites/s", "cores": 1, "kind": kind,
                               "sample": "first %d sites (%d records): %s" % (m, wn, what), "parity_records_checked": wn, "parity_bytes_checked": len(wb)}
    for b in (hp, hr, ho):
        b.free()
    return out


def host_link_probe(torch, dist, world, mb=256, reps=6):
    """What the box gives the host-buffer paths: pinned host <-> device copies with plain cudaMemcpyAsync
    (one call per copy, torch's copy_), every rank of the run at the same time: H2D alone, D2H alone, both at once on two streams.
    GB/s summed over the ranks; the e2e legs move 105 B up and 201 B down per site, so sites/s <= min(up / 105, down / 201)."""
    n = mb << 20
    hu = torch.empty(n, dtype=torch.uint8).pin_memory()
    hd = torch.empty(n, dtype=torch.uint8).pin_memory()
    du = torch.empty(n, dtype=torch.uint8, device="cuda")
    dd = torch.zeros(n, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s_up):
                    du.copy_(hu, non_blocking=True)
            if down:
                with torch.cuda.stream(s_dn):
                    hd.copy_(dd, non_blocking=True)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * reps * n / float(t.item()) / 1e9

    run(True, True)
    out = {"h2d_alone_gbs": run(True, False), "d2h_alone_gbs": run(False, True)}
    both = run(True, True)
    out.update({"h2d_and_d2h_each_gbs": both, "ranks": world, "copy_mb": mb,
                "e2e_ceiling_sites_per_s": both * 1e9 / 201.0,
                "note": "pinned host <-> device, one cudaMemcpyAsync per copy, all ranks at once; ceiling = the D2H rate with H2D running / 201 B per site"})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="bsgpu")
    ap.add_argument("--sites", type=float, default=1e9, help="resident sites per GPU (config 2: 1e9)")
    ap.add_argument("--e2e-sites", type=float, default=8e6, help="sites per e2e step (host buffers)")
    ap.add_argument("--fused-sites", type=float, default=50e6, help="window of the block-path (pileup + model) measurement")
    ap.add_argument("--bam-sites", type=float, default=50e6, help="window of the BAM-records-to-calls measurement (bsgpu_call_bam): config 3 is 50 M")
    ap.add_argument("--cpu-diff-records", type=float, default=0, help="records of that stream the CPU reference re-runs for the diff (0: all of them)")
    ap.add_argument("--deep-sites", type=float, default=10e6, help="sites of the 500x panel (config 4: 10 Mb)")
    ap.add_argument("--genome-scale", type=int, default=0, help="genome leg: hg38 contig lengths divided by this (0: by the host's free memory)")
    ap.add_argument("--genome-batch-mb", type=int, default=384, help="genome leg: batch size of the streaming session")
    ap.add_argument("--genome-sessions", type=int, default=2, help="genome leg: sessions (contexts) per GPU, regions dealt out between them")
    ap.add_argument("--no-genome", action="store_true", help="skip the genome leg")
    ap.add_argument("--binary-sites", type=float, default=50e6, help="whole-program leg: sites of the BAM file both bs_call binaries read (config 1 = configs[0] is 50 M)")
    ap.add_argument("--legs", default="e2e,block,bam,writer,genome,binary,cpu", help="secondary legs to run (comma separated)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3
    legs = set(args.legs.split(","))
    import faulthandler
    faulthandler.enable()
    if os.environ.get("BENCH_WATCHDOG_S"):          # debugging aid: dump every thread's stack if the run takes longer than this
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG_S"]), exit=True)

    import numpy as np
    import torch
    import torch.distributed as dist
    from bs_call_b200 import lib as bslib
    from bs_call_b200.records import GT_METH, PILEUP

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    numa_note = bind_to_gpu_numa_node(torch, local) if world > 1 else "single rank: not bound"
    if world > 1:
        # the reader's host stages (framer, block builder) run thread pools: share the cores between the ranks of the box
        per_rank = str(max(2, min(len(os.sched_getaffinity(0)), (os.cpu_count() or 2) // world)))
        os.environ.setdefault("BSGPU_BUILDER_THREADS", per_rank)
        os.environ.setdefault("BSGPU_FRAMER_THREADS", per_rank)
        os.environ.setdefault("BSGPU_FEED_THREADS", str(max(2, min(8, int(per_rank) // 2))))
    gpu = bslib.BsGpu(device=local)
    # a dedicated torch stream: its handle is what the C ABI launches on, and what the CUDA events below are recorded on
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    n_sites = int(args.sites)
    n_slab = 16
    slab = (n_sites + n_slab - 1) // n_slab
    slab += slab & 1                                 # even -> every slab starts 16-byte aligned (104 B records)
    d_pile = torch.empty(n_sites * 104 + 16, dtype=torch.uint8, device="cuda")
    d_ref = torch.empty(n_sites + 16, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(slab * 200 + 16, dtype=torch.uint8, device="cuda")
    d_skip = torch.empty(slab + 16, dtype=torch.uint8, device="cuda")
    first_site = rank * n_sites
    for s in range(n_slab):
        a = s * slab
        m = min(slab, n_sites - a)
        if m > 0:
            gpu.synth_sites_dev(SEED, first_site + a, m, MEAN_DEPTH, d_pile.data_ptr() + a * 104, d_ref.data_ptr() + a, stream)
    torch.cuda.synchronize()
    called_total = int((d_pile.view(torch.int32)[16:n_sites * 26:26] > 0).sum().item())      # pileup.n > 0

    def step():
        for s in range(n_slab):
            a = s * slab
            m = min(slab, n_sites - a)
            if m > 0:
                gpu.call_sites_dev(d_pile.data_ptr() + a * 104, d_ref.data_ptr() + a, m, d_out.data_ptr(), d_skip.data_ptr(), stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    gpu.guard_read(reset=True)
    g0 = gpu.stats()
    l0 = g0["kernel_launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    g1 = gpu.stats()
    launches = g1["kernel_launches"] - l0
    # guard bands (hard part 1): sites of the timed steps whose call hangs on the last bits of a likelihood
    guard = {"sites_per_step": n_sites, "near_tie_sites_per_step": (g1["near_tie_sites"] - g0["near_tie_sites"]) // args.steps,
             "exact_tie_sites_per_step": (g1["exact_tie_sites"] - g0["exact_tie_sites"]) // args.steps,
             "note": "two best genotype log-likelihoods within 1e-9 relative (near) or equal (exact): the call may differ from the CPU's at these sites only; "
                     "listed by bsgpu_guard_read, counted in bsgpu_stats"}
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tot_called = torch.tensor([called_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_called, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = float(tot_called.item()) / (ms_per_step * 1e-3)

    # spot check of the last slab against the oracle (not timed): the numbers above are for a path that is right
    parity = None
    if rank == 0:
        from oracle.bindings import Oracle
        from tests import util
        a = (n_slab - 1) * slab
        m = min(min(slab, n_sites - a), 20000)
        p = d_pile[a * 104:(a + m) * 104].cpu().numpy().view(PILEUP)
        r = d_ref[a:a + m].cpu().numpy()
        got = d_out[:m * 200].cpu().numpy().view(GT_METH)
        gskip = d_skip[:m].cpu().numpy()
        wout, wskip = Oracle().call_sites(p, r, nthreads=4)
        parity = {"sites_checked": util.assert_gt_meth_close(got, gskip, wout, wskip), "against": "oracle"}

    # ---- roofline of the dominant kernel (k_call_sites): per launch, all sites of the slab move 306 B each
    peak, peak_kind = measured_peaks()
    launch_ms = ms / (args.steps * n_slab)
    achieved = BYTES_PER_SITE * slab / (launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if "k_call_sites_bench" in tj and abs(tj["k_call_sites_bench"]["sites_per_launch"] - slab) <= 2:
                traffic = tj["k_call_sites_bench"]["dram_bytes_per_launch"]          # measured on a launch of this very size
            else:
                traffic = tj["k_call_sites"]["dram_bytes_per_site"] * slab
        except Exception:
            traffic = None
    roofline = {"kernel": "k_call_sites", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                "algorithmic_bytes_per_site": BYTES_PER_SITE, "sites_per_launch": slab, "launch_ms": launch_ms}

    # ---- e2e: the same metric through the C ABI with host buffers (H2D + kernel + D2H inside the timed region)
    n_e2e = int(min(args.e2e_sites, n_sites))
    hp = bslib.HostBuffer(n_e2e, PILEUP)
    hr = bslib.HostBuffer(n_e2e, np.uint8)
    ho = bslib.HostBuffer(n_e2e, GT_METH)
    hs = bslib.HostBuffer(n_e2e, np.uint8)
    hp.array.view(np.uint8)[:] = d_pile[:n_e2e * 104].cpu().numpy()
    hr.array[:] = d_ref[:n_e2e].cpu().numpy()
    e2e_called = int((hp.array["n"] > 0).sum())
    for _ in range(2):
        gpu.call_sites(hp.array, hr.array, out=ho.array, skip=hs.array)
    barrier()
    e2e_steps = max(args.steps, 5)                  # a call is ~30 ms: at least five of them, whatever K is
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        gpu.call_sites(hp.array, hr.array, out=ho.array, skip=hs.array)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    ce = torch.tensor([e2e_called], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(ce, op=dist.ReduceOp.SUM)
    e2e = {"value": float(ce.item()) * e2e_steps / float(te.item()), "unit": "sites/s", "h2d_bytes_per_step": n_e2e * 105,
           "d2h_bytes_per_step": n_e2e * 201, "sites_per_step": n_e2e,
           "calls_timed": e2e_steps,
           "note": "bsgpu_call_sites on pinned host arrays; 256 Ki-site chunks ping-pong on two streams"}
    assert (hs.array == (hp.array["n"] == 0)).all()
    for b in (hp, hr, ho, hs):
        b.free()
    e2e["host_link"] = host_link_probe(torch, dist, world)

    # ---- block path (segments -> pileup -> model) on a synthetic 30x window (config 3 shape), device resident
    block = None
    del d_out, d_skip
    torch.cuda.empty_cache()
    log("headline %.3g sites/s, e2e %.3g sites/s; block path ..." % (value, e2e["value"]))
    try:
        if "block" not in legs:
            raise RuntimeError("leg skipped (--legs)")
        fsz = int(args.fused_sites)
        L, depth = 150, 30.0
        ns = gpu.synth_block_nseg(fsz, L, depth)
        d_seg = torch.empty(ns * 16 + 16, dtype=torch.uint8, device="cuda")
        d_b = torch.empty(ns * L + 16, dtype=torch.uint8, device="cuda")
        d_r = torch.empty(fsz + 16, dtype=torch.uint8, device="cuda")
        d_v = torch.empty(fsz * 208 + 16, dtype=torch.uint8, device="cuda")
        d_p = torch.empty(fsz * 104 + 16, dtype=torch.uint8, device="cuda")
        gpu.synth_block_dev(SEED + rank, 1000, fsz, L, depth, d_seg.data_ptr(), ns, d_b.data_ptr(), ns * L, d_r.data_ptr(), stream)

        def timed(fn):
            for _ in range(3):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                fn()
            b.record()
            barrier()
            return a.elapsed_time(b) / args.steps

        bms = timed(lambda: gpu.call_block_dev(d_seg.data_ptr(), ns, d_b.data_ptr(), d_r.data_ptr(), 1000, fsz, d_v.data_ptr(), stream))
        fcalled = int((d_v.view(torch.uint8)[201::208][:fsz] == 0).sum().item())
        pms = timed(lambda: gpu.pileup_block_dev(d_seg.data_ptr(), ns, d_b.data_ptr(), 1000, fsz, d_p.data_ptr(), stream))
        os.environ["BSGPU_FUSED"] = "1"
        gfused = bslib.BsGpu(device=local)
        del os.environ["BSGPU_FUSED"]
        fms = timed(lambda: gfused.call_block_dev(d_seg.data_ptr(), ns, d_b.data_ptr(), d_r.data_ptr(), 1000, fsz, d_v.data_ptr(), stream))
        gfused.close()
        # the gather of part k + 1 on its own stream next to the model of part k (BSGPU_BLOCK_OVERLAP): same kernels, same bytes
        d_v2 = torch.empty(fsz * 208 + 16, dtype=torch.uint8, device="cuda")
        os.environ["BSGPU_BLOCK_OVERLAP"] = "0" if BLOCK_OVERLAP_DEFAULT else "1"
        galt = bslib.BsGpu(device=local)
        del os.environ["BSGPU_BLOCK_OVERLAP"]
        oms = timed(lambda: galt.call_block_dev(d_seg.data_ptr(), ns, d_b.data_ptr(), d_r.data_ptr(), 1000, fsz, d_v2.data_ptr(), stream))
        torch.cuda.synchronize()
        alt_same = bool(torch.equal(d_v[:fsz * 208], d_v2[:fsz * 208]))
        galt.close()
        del d_v2
        in_bytes = ns * (L + 16)

        def roof(nbytes, ms_):
            return {"bound": "hbm", "achieved": nbytes / (ms_ * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": nbytes / (ms_ * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_site": nbytes / fsz}

        block = {"workload": "synthetic %dx %d-bp bisulfite reads over a %d-site window (config 3 shape), device resident" % (int(depth), L, fsz),
                 "sites_called": fcalled,
                 "default": {"kernels": "k_bin_* + k_pileup_tile<pileup> -> pileup[] in HBM -> k_call_sites<vcf>", "ms": bms,
                             "sites_per_s": fcalled / (bms * 1e-3), "roofline": dict(roof(in_bytes + fsz * (1 + 208), bms),
                                 note="algorithmic bytes = SURVEY.md 8d fused figure (segments + bases + ref in, gt_vcf out); the pileup[] scratch between the two kernels "
                                      "(104 B/site written and read back) is this implementation's own traffic and is not counted")},
                 "pileup_only": {"kernels": "k_bin_* + k_pileup_tile<pileup>", "ms": pms, "sites_per_s": fcalled / (pms * 1e-3),
                                 "roofline": roof(in_bytes + fsz * 104, pms)},
                 "fused_variant": {"kernels": "k_bin_* + k_pileup_tile<fused> (BSGPU_FUSED=1)", "ms": fms, "sites_per_s": fcalled / (fms * 1e-3),
                                   "roofline": roof(in_bytes + fsz * (1 + 208), fms)},
                 ("serial_variant" if BLOCK_OVERLAP_DEFAULT else "overlapped_variant"): {
                     "kernels": "the same two kernels, " + ("one after the other on one stream (BSGPU_BLOCK_OVERLAP=0)" if BLOCK_OVERLAP_DEFAULT else
                                                            "window in parts of 2 Mi sites, gather of part k + 1 on its own stream next to the model of part k (BSGPU_BLOCK_OVERLAP=1)"),
                     "ms": oms, "sites_per_s": fcalled / (oms * 1e-3), "roofline": roof(in_bytes + fsz * (1 + 208), oms),
                     "gt_vcf_bytes_identical_to_default": alt_same}}
        # deep targeted panel (config 4: 500x single-end 150-bp reads over 10 Mb), the stress case for pileup accumulation.  One
        # window holds at most 4 Gi bases (32-bit offsets into bases[]), so the 10 Mb go through as windows of 2.5 Mb that
        # share the device buffers; times are summed over the windows.
        try:
            ddepth = 500.0
            dtot = int(min(args.deep_sites, 10_000_000))
            dwin = int(min(dtot, 2_500_000))
            nwin = (dtot + dwin - 1) // dwin
            dns = gpu.synth_block_nseg(dwin, L, ddepth)
            dd_seg = torch.empty(dns * 16 + 16, dtype=torch.uint8, device="cuda")
            dd_b = torch.empty(dns * L + 16, dtype=torch.uint8, device="cuda")
            dpms = dbms = 0.0
            dcalled = 0
            for wi in range(nwin):
                gpu.synth_block_dev(SEED + 99 + rank + 7 * wi, 1000, dwin, L, ddepth, dd_seg.data_ptr(), dns, dd_b.data_ptr(), dns * L, d_r.data_ptr(), stream)
                dpms += timed(lambda: gpu.pileup_block_dev(dd_seg.data_ptr(), dns, dd_b.data_ptr(), 1000, dwin, d_p.data_ptr(), stream))
                dbms += timed(lambda: gpu.call_block_dev(dd_seg.data_ptr(), dns, dd_b.data_ptr(), d_r.data_ptr(), 1000, dwin, d_v.data_ptr(), stream))
                dcalled += int((d_v.view(torch.uint8)[201::208][:dwin] == 0).sum().item())
            dsz = dwin * nwin
            dbytes = nwin * (dns * (L + 16) + dwin * 104)
            block["deep_panel"] = {"workload": "synthetic %dx %d-bp reads over %d sites in %d windows (config 4: 10 Mb panel), device resident" % (int(ddepth), L, dsz, nwin),
                                   "sites_called": dcalled,
                                   "pileup_only": {"ms": dpms, "sites_per_s": dsz / (dpms * 1e-3), "bases_per_s": nwin * dns * L / (dpms * 1e-3),
                                                   "roofline": {"bound": "hbm", "achieved": dbytes / (dpms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                                                "frac": dbytes / (dpms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_site": dbytes / dsz}},
                                   "default": {"ms": dbms, "sites_per_s": dsz / (dbms * 1e-3),
                                               "roofline": {"bound": "hbm", "achieved": (dbytes + dsz * 105) / (dbms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                                            "frac": (dbytes + dsz * 105) / (dbms * 1e-3) / 1e9 / peak,
                                                            "algorithmic_bytes_per_site": (dbytes + dsz * 105) / dsz}},
                                   "envelope_overflow_sites": gpu.stats()["qsum_overflow"]}
            del dd_seg, dd_b
        except Exception as e:
            block["deep_panel"] = {"error": repr(e)}
        # end to end through the host-buffer entry point (what the drop-in calls): pinned host segments / bases / ref in,
        # gt_vcf[] out, H2D + binning + pileup + model + D2H inside the timed region; a 16 Mi-site window of the same data
        from bs_call_b200.records import GT_VCF, SEG, TEMPLATE
        esz = int(min(fsz, 16 * 1024 * 1024))
        ens = gpu.synth_block_nseg(esz, L, depth)
        gpu.synth_block_dev(SEED + rank, 1000, esz, L, depth, d_seg.data_ptr(), ens, d_b.data_ptr(), ens * L, d_r.data_ptr(), stream)
        torch.cuda.synchronize()
        hseg = bslib.HostBuffer(ens, SEG)
        hb = bslib.HostBuffer(ens * L, np.uint8)
        hr = bslib.HostBuffer(esz, np.uint8)
        hv = bslib.HostBuffer(esz, GT_VCF)
        hseg.array.view(np.uint8)[:] = d_seg[:ens * 16].cpu().numpy()
        hb.array[:] = d_b[:ens * L].cpu().numpy()
        hr.array[:] = d_r[:esz].cpu().numpy()
        for _ in range(2):
            gpu.call_block(hseg.array, hb.array, hr.array, 1000, esz, out=hv.array)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            gpu.call_block(hseg.array, hb.array, hr.array, 1000, esz, out=hv.array)
        torch.cuda.synchronize()
        es = (time.perf_counter() - t0) / args.steps
        ecalled = int((hv.array["skip"] == 0).sum())
        block["e2e"] = {"value": ecalled / es, "unit": "sites/s", "sites_per_step": esz, "h2d_bytes_per_step": ens * (16 + L) + esz,
                        "d2h_bytes_per_step": esz * 208, "note": "bsgpu_call_block on pinned host arrays (what the drop-in call_genotypes_ML calls)"}
        # the reference's own call_genotypes_ML (serial pileup + calc threads on all cores) on a bounded slice of that block
        if rank == 0 and world == 1 and not args.no_cpu:
            from oracle.bindings import Reference, reference_available
            csz = 1_500_000
            segs_h = hseg.array
            m = (segs_h["pos"] >= 1000) & (segs_h["pos"] + segs_h["len"] <= 1000 + csz)
            sub = segs_h[m]
            T = np.zeros(len(sub), dtype=TEMPLATE)
            T["forward_position"] = sub["pos"]
            T["read_off"][:, 0] = sub["off"]
            T["read_len"][:, 0] = sub["len"]
            T["present"][:, 0] = 1
            T["mapq"][:, 0] = sub["mapq"]
            T["orientation"] = sub["flags"] & 1
            T["bs_strand"] = (sub["flags"] >> 1) & 3
            refc = np.concatenate([hr.array[:csz], np.zeros(2, np.uint8)])
            ncores = os.cpu_count() or 1
            if reference_available():
                R = Reference(calc_threads=ncores)
                R.call_block(T[:2000], hb.array, refc, 1000, 1000 + csz - 1)          # warm-up (allocations)
                t0 = time.perf_counter()
                pile_c, vcf_c = R.call_block(T, hb.array, refc, 1000, 1000 + csz - 1)
                cs = R.last_call_seconds()          # call_genotypes_ML until all sites ready; harness set-up excluded
                kind, what = "reference", "reference call_genotypes_ML + call_thread (oracle/_ref/libbsref.so), harness threads as built"
            else:
                from oracle.bindings import Oracle
                o = Oracle()
                t0 = time.perf_counter()
                pile_c = o.pileup_block(T, hb.array, 1000, 1000 + csz - 1)
                out_c, skip_c = o.call_sites(pile_c, refc[:csz], nthreads=ncores)
                cs = time.perf_counter() - t0
                vcf_c = None
                kind, what = "port", "oracle pileup (1 thread, as the reference) + model on all cores"
            ccalled = int((pile_c["n"] > 0).sum())
            block["cpu_baseline"] = {"value": ccalled / cs, "unit": "sites/s", "cores": ncores, "kind": kind,
                                     "sample": "%d sites (%d called, %d segments) of the same block; %s" % (csz, ccalled, len(sub), what)}
            if vcf_c is not None:
                from tests import util
                block["cpu_baseline"]["parity_sites_checked"] = util.assert_vcf_close(hv.array[:csz - 300], vcf_c[:csz - 300])
        for hb_ in (hseg, hb, hr, hv):
            hb_.free()
        del d_seg, d_b, d_r, d_v, d_p
    except Exception as e:            # reported, never hidden
        import traceback
        block = {"error": repr(e), "trace": traceback.format_exc()[-800:]}

    # ---- the whole path from BAM records: decode -> blocks -> normalise -> pileup -> model (bsgpu_call_bam), host buffers
    bam = None
    log("BAM records leg (config 3) ...")
    try:
        if "bam" not in legs:
            raise RuntimeError("leg skipped (--legs)")
        bam = bam_path(args, gpu, bslib, torch, np, stream, rank, world, local)
    except Exception as e:
        import traceback
        bam = {"error": repr(e), "trace": traceback.format_exc()[-800:]}

    # ---- the writer's derivations on the device: gt_vcf[] -> BCF records; count vectors -> BCF records end to end
    writer = None
    log("writer leg ...")
    try:
        del d_pile, d_ref
        torch.cuda.empty_cache()
        if "writer" not in legs:
            raise RuntimeError("leg skipped (--legs)")
        writer = writer_path(args, gpu, bslib, torch, np, stream, rank, world, local)
    except Exception as e:
        import traceback
        writer = {"error": repr(e), "trace": traceback.format_exc()[-800:]}

    # ---- the genome: 24 contigs sharded over the ranks, streaming sessions, ordered merge (strong scaling)
    genome = None
    log("genome leg (config 5 shape) ...")
    if not args.no_genome and "genome" in legs:
        try:
            torch.cuda.empty_cache()
            genome = genome_path(args, gpu, bslib, torch, np, stream, rank, world, local)
        except Exception as e:
            import traceback
            genome = {"error": repr(e), "trace": traceback.format_exc()[-1200:]}

    # ---- the whole program: the reference's bs_call binary against the same program with the seam files, from files
    fullbin = None
    if rank == 0 and world == 1 and "binary" in legs:
        log("whole-program leg (config 1 shape) ...")
        try:
            torch.cuda.empty_cache()
            fullbin = full_binary_path(args, gpu, bslib, torch, np, stream, local)
        except Exception as e:
            import traceback
            fullbin = {"error": repr(e), "trace": traceback.format_exc()[-1200:]}

    cpu = None
    log("cpu baseline of the headline ...")
    if rank == 0 and world == 1 and not args.no_cpu and "cpu" in legs:
        nthreads = os.cpu_count() or 1
        n_cpu = calibrate_cpu_sample(nthreads, 12.0)
        dt, kind, sample, called, _ = cpu_path(n_cpu, nthreads)
        cpu = {"value": called / dt, "unit": "sites/s", "cores": nthreads, "kind": kind, "sample": sample}

    if rank == 0:
        line = {"metric": "genome sites called/sec", "value": value, "unit": "sites/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "BASELINE.json configs[1]: likelihood microbench, 1e9 synthetic per-site count vectors per GPU",
                           "sites_per_gpu": n_sites, "sites_called_per_gpu": called_total, "mean_depth": MEAN_DEPTH, "seed": SEED,
                           "parallelism": "sites sharded over %d rank(s), no collective" % world, "host": numa_note,
                           "l2": "inputs larger than L2: each step streams %.1f GB of distinct records" % (n_sites * 105 / 1e9)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "block_path": block, "bam_path": bam, "writer_path": writer, "genome_path": genome, "full_binary": fullbin, "guard_bands": guard, "parity_spot_check": parity}
        print(json.dumps(line))
    gpu.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
