#!/usr/bin/env python
"""bench.py — throughput of the RX synchronisation hot path (SyncwordDetection) on B200.

Metric (BASELINE.json): complex Msps (cf32) through RX sync; % of the HBM roofline; the
reference's CPU algorithm timed on the same box's host cores beside it.

Workload (config.workload): BASELINE.json configs[1] — syncword detection over a 2^30-sample
synthetic cf32 capture per GPU (QPSK, 4 sps, RRC, Es/N0 20 dB, CFO 0.005 rad/sample), K = 9
frequency hypotheses (min/max_freq_bin = -/+4), power_threshold 9.5, time_threshold 768.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log2n 30] [--bins 4]
  torchrun ... bench.py --gpus N ...      one rank per GPU; time shards + halo, no data-path collective
  python bench.py --impl reference ...    the reference's CPU block code (oracle/_ref, else the oracle port) on the host cores
  python bench.py --workload chain        configs[2]: front end -> detection -> CFC + SymbolFilter -> wipe-off + Costas
  python bench.py --workload channels     configs[4] channel mode (under torchrun: channels partitioned over ranks)
  python bench.py --bins 16 --esn0 0      configs[3]: low-SNR, K = 33 hypotheses (sweep: scripts/threshold_sweep.py)

The capture is written into HBM by the library's own generator (csrc/stimulus.cu, index-pure: every rank
generates its shard + halo of the same endless stream).  A step is one pass of the whole hot path under the
block contract — correlator with the fused delayed output (16 B/sample: 8 read + 8 written,
PM/syncword_detection.hpp:318-319), peak detector, refine, records to host — over the device-resident
capture.  The capture (8 GiB at 2^30) is far larger than L2 (126 MB), so no L2 flush is needed between steps.
At N = 1 the line also carries compact records of the other BASELINE configurations (`configs`): K = 1 (the
HBM-bound case), configs[2] chain, configs[3] K = 33 at 0 dB, configs[4] 64-channel mode.
Multi-rank runs exchange only the shards' (T+1)-entry chain tables, as host bytes over gloo: no NCCL on the data path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FFT, TAU = 2048, 768
BYTES_PER_SAMPLE = 16  # block contract, 8 B read + 8 B written (delayed output); detection-only = 8


def flop_per_sample(K: int, S: int = 1752) -> float:
    """nominal 5 N log2 N per FFT, 6 per complex multiply, 3 per |.|^2."""
    return ((1 + K) * 112640 + 9 * K * 2048 + 3 * 1024) / S


class ClockSampler:
    """Streams `nvidia-smi -lms 100` (B200_PROFILING.md clocks line).  Started BEFORE the warm-up so that the
    tool is already streaming when the timed region begins; every row is stamped on arrival and only rows
    that arrived inside [mark_begin(), stop()] count.  A timed region shorter than the sampling period can
    see no row at all: then the rows of the warm-up (same kernels, same load) are used and `window` says so."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.rows, self.t_begin, self.thread = index, None, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def stop(self) -> dict:
        t_end = time.perf_counter()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            if self.thread is not None:
                self.thread.join(timeout=2)
        ok = [(t, r) for t, r in self.rows if len(r) >= 7]
        t0 = self.t_begin if self.t_begin is not None else 0.0
        inside = [r for t, r in ok if t0 <= t <= t_end + 0.15]
        window = "timed region"
        if not inside:
            inside, window = [r for _, r in ok][-5:], "warm-up (timed region shorter than the sampling period)"

        def num(v):
            try:
                return float(v)
            except ValueError:
                return None
        sm = [num(r[0]) for r in inside if num(r[0]) is not None]
        mx = [num(r[1]) for r in inside if num(r[1]) is not None]
        pw = [num(r[2]) for r in inside if num(r[2]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm), "window": window}


def workload_config(log2n: int, K: int, esn0: float = 20.0, thr: float = 9.5) -> dict:
    """The `config` object of the JSON line: the same for the B200 arm and the reference arm."""
    which = "configs[1]" if (K == 9 and esn0 == 20.0) else "configs[3] (low-SNR / wide-CFO search)" if K > 9 else "configs[1] variant"
    return {"workload": f"BASELINE {which}: syncword detection over a 2^{log2n}-sample synthetic cf32 "
                        f"capture per GPU, K={K} hypotheses, QPSK 4 sps RRC, Es/N0 {esn0:g} dB, CFO 0.005 rad/sample",
            "samples_per_gpu": 1 << log2n, "fft_size": FFT, "time_threshold": TAU, "power_threshold": thr,
            "l2": "inputs (8 B/sample resident capture) larger than L2; no flush needed",
            "contract": "block contract: delayed output span written (16 B/sample); the reference arm does the same",
            "sharding": "contiguous time shards + 1-block halo; (T+1)-entry chain tables exchanged as host bytes "
                        "(shared memory between the ranks of the box, gloo as fallback); no NCCL on the data path"}


def measured_traffic(log2n: int, K: int, contract: str = "block", kernel: str = "correlate_kernel"):
    """DRAM bytes (read + write) of one launch from the committed `ncu --set full` capture of this configuration
    (profiles/r2_traffic.json: {"<log2n>/<K>/<contract>": {kernel: {...}}}), or None when none is committed."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        k = t[f"{log2n}/{K}/{contract}"][kernel]
        return k["dram_read_bytes"] + k["dram_write_bytes"]
    except Exception:
        return None


def binding_roof(K: int) -> tuple:
    """What ncu says binds correlate_kernel (profiles/r2_ncu_summary_*.md): (field value, sentence)."""
    if K <= 1:
        return "hbm+fp32", ("K=1: 141 nominal flop/sample against 8-16 B/sample — the one configuration where HBM "
                            "matters (SURVEY §8d); see profiles/r2_ncu_summary_k1.md for the measured split")
    return "fp32_pipe", (f"K={K}: {flop_per_sample(K):.0f} nominal flop/sample, far above the ridge — bound by the FP32 "
                         "pipe (packed FFMA2/FADD2 occupy it two passes each; ncu: math-pipe-throttle is the top "
                         "stall, LSU wavefronts 71 %, DRAM 15 %), not by HBM")


def measured_fp32_peak():
    """FP32 FMA peak measured on this pool's B200 (SURVEY §8d asks for a measured figure), or None."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "r1_fp32_peak.json")))["fp32_fma_tflops"])
    except Exception:
        return None


def rx_settings(bins: int, thr: float = 9.5):
    from gr4_packet_modem_b200.firdes import BPSK, SYNCWORD, unit_energy_rrc

    return dict(rrc_taps=unit_energy_rrc(), syncword=SYNCWORD, constellation=BPSK, min_freq_bin=-bins,
                max_freq_bin=bins, time_threshold=TAU, power_threshold=thr)


CPU_KIND_NOTE = {
    "reference": "the reference's own block code (PM/syncword_detection.hpp compiled unmodified from the reference "
                 "tree against a stand-in GR4 runtime, oracle/_ref/librefblocks.so) with a radix-2 FFT standing in "
                 "for FFTW 3.3.10, which is not in the tree (DESIGN.md §7)",
    "port": "oracle port of PM/syncword_detection.hpp (bit-identical to the reference's block code), radix-2 FFT "
            "in place of FFTW",
}


def cpu_reference_kind(force_port: bool = False) -> str:
    """"reference": oracle/_ref/librefblocks.so exists — the reference's own SyncwordDetection block code
    (PM/syncword_detection.hpp compiled unmodified, stand-in GR4 runtime; FFT = radix-2 stand-in for the
    absent FFTW 3.3.10); "port": the oracle's restatement (bit-identical results, same FFT)."""
    if force_port:
        return "port"
    try:
        from oracle import refblocks as rb

        return "reference" if rb.available() else "port"
    except Exception:
        return "port"


def cpu_reference_rate(bins: int, samples_per_thread: int, threads: int, steps: int, warmup: int,
                       esn0: float = 20.0, thr: float = 9.5, kind: str = "port"):
    """The reference's CPU implementation of the path: `threads` independent streams, one per host thread (one
    GR4 block instance runs on one worker thread), fed 65536-item chunks like the GR4 ring does.  kind:
    see cpu_reference_kind().  Returns (aggregate Msps, seconds per step)."""
    from gr4_packet_modem_b200.stimulus import packet_capture
    from oracle import pyoracle as po

    po.build(ref=False)
    x, _ = packet_capture(samples_per_thread, seed=1, esn0_db=esn0, cfo=0.005)
    s = rx_settings(bins, thr)
    if kind == "reference":
        from oracle import refblocks as rb

        sds = [rb.SyncwordDetection(s["rrc_taps"], s["syncword"], s["constellation"], -bins, bins, TAU, thr)
               for _ in range(threads)]
    else:
        sds = [po.SyncwordDetection(s["rrc_taps"], s["syncword"], s["constellation"], -bins, bins, TAU, thr,
                                    fft_kind=po.FFT_RADIX2) for _ in range(threads)]
    consumed = [0] * threads

    def work(i):
        if kind == "reference":
            c, _, _ = sds[i].run(x, chunk=65536)
        else:
            c, _, _ = sds[i].run(x, chunk=65536, want_output=True)
        consumed[i] = c

    def one_step():
        ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    for _ in range(warmup):
        one_step()
    dt = [one_step() for _ in range(steps)]
    total = sum(consumed)
    sec = sum(dt) / len(dt)
    return total / sec / 1e6, sec


def zero_input_rate(bins: int, kind: str):
    """benchmarks/benchmark_syncword_detection.cpp in miniature: 2^22 zeros through one block instance."""
    try:
        from oracle import pyoracle as po

        s = rx_settings(bins)
        x = np.zeros(1 << 22, np.complex64)
        if kind == "reference":
            from oracle import refblocks as rb

            blk = rb.SyncwordDetection(s["rrc_taps"], s["syncword"], s["constellation"], -bins, bins, TAU, 9.5)
            t0 = time.perf_counter()
            c, _, _ = blk.run(x, chunk=65536)
        else:
            blk = po.SyncwordDetection(s["rrc_taps"], s["syncword"], s["constellation"], -bins, bins, TAU, 9.5,
                                       fft_kind=po.FFT_RADIX2)
            t0 = time.perf_counter()
            c, _, _ = blk.run(x, chunk=65536, want_output=True)
        return c / (time.perf_counter() - t0) / 1e6
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_thread = 1 << 22   # the same bounded sample as the B200 arm's cpu_baseline leg
    kind = cpu_reference_kind()
    rate, sec = cpu_reference_rate(args.bins, per_thread, threads, max(args.steps, 1), max(args.warmup, 1),
                                   args.esn0, args.thr, kind)
    K = 2 * args.bins + 1
    line = {
        "impl": "reference", "metric": "complex Msps (cf32) through RX sync (SyncwordDetection)",
        "value": rate, "unit": "Msps", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.log2n, K, args.esn0, args.thr),
        "cpu_baseline": {"value": rate, "unit": "Msps", "cores": threads, "kind": kind,
                         "sample": f"bounded sample of the workload (config.samples_per_gpu names the B200 arm's "
                                   f"capture, not this sample): {threads} independent streams x 2^22 samples of the same "
                                   "signal model per step, one per host thread, delayed output produced; "
                                   + CPU_KIND_NOTE[kind]},
        "e2e": {"value": rate, "unit": "Msps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "zero_input_single_core_msps": zero_input_rate(args.bins, kind),
        "zero_input_note": "BASELINE configs[0]: the reference's own benchmark feeds zeros (NullSource -> Head -> "
                           "SyncwordDetection -> NullSink, benchmarks/benchmark_syncword_detection.cpp:28-70); one "
                           "stream, one core, same block code as above.  benchmarks/results.md:35-41 publishes for it, with "
                           "FFTW on a Ryzen 7 5800X: 49-51 / 29 / 20-21 / 16 / 13 Msps at 0 / 1 / 2 / 3 / 4 frequency "
                           "bins — i.e. the radix-2 stand-in FFT on this host core understates the real reference by "
                           "the ratio of that figure to this one",
        "published_zero_input_msps": {"0": 50.0, "1": 29.0, "2": 20.5, "3": 16.0, "4": 13.0}.get(str(args.bins)),
    }
    pub, zi = line["published_zero_input_msps"], line["zero_input_single_core_msps"]
    line["reference_understated_by"] = (pub / zi) if (pub and zi) else None
    line["reference_understated_note"] = ("published FFTW-based single-core rate / this box's single-core rate with the "
                                          "radix-2 stand-in FFT: multiply `value` by it for what the real reference "
                                          "(FFTW 3.3.10) would reach on comparable cores")
    print(json.dumps(line))


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def hbm_roofline(kernel: str, alg_bytes: float, ms: float, peaks: dict, note: str, bound: str = "hbm") -> dict:
    """`achieved`/`peak`/`frac` are always the HBM figures the contract asks for (algorithmic bytes / device time /
    measured copy bandwidth); `bound` names the roof that actually binds the kernel."""
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": bound, "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": None, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms,
            "peak_source": "MEASURED_PEAKS.json (burst copy)" if peaks else "fallback 6650 GB/s", "note": note}


def cpu_chain_rate(samples_per_thread: int, threads: int, bins: int, esn0: float, thr: float):
    """configs[2] on the host cores: the reference's own classes chained the way apps/packet_transceiver.cpp and
    PM/packet_receiver.hpp wire them — PfbArbResampler -> Rotator -> SyncwordDetection -> CoarseFrequencyCorrection ->
    SymbolFilter -> SyncwordWipeoff -> CostasLoop — one chain per host thread.  -> (aggregate Msps, kind) or None."""
    try:
        from gr4_packet_modem_b200.firdes import SYNCWORD, lowpass_prototype_taps, pfb_matched_filter_taps
        from gr4_packet_modem_b200.stimulus import packet_capture
        from oracle import refblocks as rb

        if not rb.available():
            return None
        x, _ = packet_capture(samples_per_thread, seed=1, esn0_db=esn0, cfo=0.0)
        s = rx_settings(bins, thr)
        rate = float(np.float32(1.0) + np.float32(1e-6) * np.float32(1.2))
        fe_taps, sf_taps = lowpass_prototype_taps(32, 40), pfb_matched_filter_taps()
        sw = np.where(np.asarray(SYNCWORD) != 0, -1.0, 1.0).astype(np.float32)
        consumed = [0] * threads

        def work(i):
            c_in, y = rb.PfbArbResampler(rate, fe_taps, 32).run(x, chunk=65536, out_chunk=70000)
            y = rb.rotator(y, 0.005)
            c, delayed, tags = rb.SyncwordDetection(s["rrc_taps"], s["syncword"], s["constellation"], -bins, bins, TAU,
                                                    thr).run(y, chunk=65536)
            corrected = rb.CoarseFrequencyCorrection(26).run(delayed, [(t.index, t.freq) for t in tags])
            sym, otags = rb.SymbolFilter(sf_taps, 32, 4, delay=44).run(corrected, [(t.index, t) for t in tags], chunk=65536)
            wiped = rb.SyncwordWipeoff(sw).run(sym, [i for i, _ in otags])
            rb.CostasLoop(0.01, "BPSK").run(wiped, [(i, q.phase) for i, q in otags])
            consumed[i] = c_in

        ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return sum(consumed) / (time.perf_counter() - t0) / 1e6
    except Exception as e:  # the baseline is a report, never a reason to lose the line
        sys.stderr.write(f"cpu_chain_rate: {e!r}\n")
        return None


def measure_chain(args, log2n: int, steps: int, warmup: int, e2e: bool, cpu: bool) -> dict:
    """BASELINE configs[2]: raw stream -> fused [PfbArbResampler(1 + 1.2 ppm) + Rotator(0.005)] ->
    SyncwordDetection (block contract: delayed pass-through + tags) -> CFC + SymbolFilter -> wipe-off + Costas, all
    device-resident.  One step = the stages over the whole capture; stage times by CUDA events on the launching stream."""
    import torch

    from gr4_packet_modem_b200 import CostasLoop, FrontEnd, SymbolFilter, SyncwordDetection, _native
    from gr4_packet_modem_b200.blocks import stream_tags_from_detection
    from gr4_packet_modem_b200.firdes import SYNCWORD, lowpass_prototype_taps, pfb_matched_filter_taps
    from gr4_packet_modem_b200.stimulus import DeviceStimulus

    dev = torch.device("cuda", 0)
    n = 1 << log2n
    K = 2 * args.bins + 1
    lib = _native.lib()
    stream = torch.cuda.current_stream().cuda_stream
    raw = DeviceStimulus(seed=1, esn0_db=args.esn0, cfo=0.0).generate(n, dev)  # native generator, csrc/stimulus.cu
    rate = float(np.float32(1.0) + np.float32(1e-6) * np.float32(1.2))
    fe_taps = lowpass_prototype_taps(32, 40)
    sf_taps = pfb_matched_filter_taps()
    sd = SyncwordDetection(**rx_settings(args.bins, args.thr), device=0)
    fe = FrontEnd(rate=rate, taps=fe_taps, phase_incr=0.005)
    fe_fma = FrontEnd(rate=rate, taps=fe_taps, phase_incr=0.005, fp_contract=True)   # opt-in mode, timed beside the default
    # PM/packet_receiver.hpp:94-115: CoarseFrequencyCorrection(delay 26) -> SymbolFilter(delay 44), fused
    sf = SymbolFilter(sf_taps, 32, 4, delay=44, fused_cfc_delay=None if args.no_cfc else 26)
    n_y = fe.max_output(n)
    y = torch.empty(n_y, dtype=torch.complex64, device=dev)        # conditioned stream
    dl = torch.empty(n_y, dtype=torch.complex64, device=dev)       # SyncwordDetection's delayed output
    # one symbol per 4 samples, plus at most one extra per syncword tag (PM/symbol_filter.hpp:160-189); with that much room
    # the bulk call may pipeline its host replay against the kernel (it cannot fail half way)
    sym = torch.empty(n_y // 4 + n_y // (TAU + 1) + 1024, dtype=torch.complex64, device=dev)
    # PM/packet_receiver.hpp:117-125, 203-214: SyncwordWipeoff(bipolar syncword) fused into CostasLoop (defaults)
    cl = CostasLoop(0.01, "BPSK")
    cl.fuse_wipeoff(np.where(np.asarray(SYNCWORD) != 0, -1.0, 1.0).astype(np.float32))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

    def step(timed=None, src_ptr=None):
        fe.restart()   # start(): fresh streaming state per pass, allocations kept
        sf.restart()
        ev[0].record()
        c_in, n_out = fe.process_device(src_ptr or raw.data_ptr(), n, y.data_ptr(), n_y, stream)
        ev[1].record()
        consumed, recs, tags = sd.detect_device(y.data_ptr(), n_out, stream, d_out_ptr=dl.data_ptr())
        ev[2].record()
        st = stream_tags_from_detection(tags)
        c_sf, n_sym, otags = sf.process_device(dl.data_ptr(), consumed, sym.data_ptr(), sym.numel(), st, stream)
        ev[3].record()
        cl.start()
        cl.process_device(sym.data_ptr(), n_sym, sym.data_ptr(), otags, stream)  # in place
        ev[4].record()
        torch.cuda.synchronize()
        if timed is not None:
            timed.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]),
                          ev[3].elapsed_time(ev[4])))
        return c_in, n_out, consumed, len(recs), n_sym, len(otags)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    launches0 = lib.b200sync_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timed = []
    e0.record()
    for _ in range(steps):
        c_in, n_out, consumed, ndet, n_sym, ntags = step(timed)
    e1.record()
    torch.cuda.synchronize()
    launches = lib.b200sync_launch_count() - launches0
    ms_per_step = e0.elapsed_time(e1) / steps
    fe_ms = statistics.mean(t[0] for t in timed)
    sd_ms = statistics.mean(t[1] for t in timed)
    sf_ms = statistics.mean(t[2] for t in timed)
    cl_ms = statistics.mean(t[3] for t in timed)
    # the opt-in fused-multiply-add front end (b200sync_fe_config::fp_contract), the stage alone, same input and buffers
    fma_ms = []
    for i in range(warmup + steps):
        fe_fma.restart()
        ev[0].record()
        fe_fma.process_device(raw.data_ptr(), n, y.data_ptr(), n_y, stream)
        ev[1].record()
        torch.cuda.synchronize()
        if i >= warmup:
            fma_ms.append(ev[0].elapsed_time(ev[1]))
    fe_fma_ms = statistics.mean(fma_ms)
    peaks = load_peaks()
    rec = {
        "metric": "complex Msps (cf32) through RX sync (fused front end + SyncwordDetection + "
                  + ("" if args.no_cfc else "CoarseFrequencyCorrection + ") + "SymbolFilter + SyncwordWipeoff + CostasLoop)",
        "value": c_in / (ms_per_step * 1e-3) / 1e6, "unit": "Msps", "ms_per_step": ms_per_step, "steps": steps,
        "config": {"workload": f"BASELINE configs[2]: fused RX front end (PfbArbResampler 1+1.2ppm + Rotator 0.005) -> "
                               f"SyncwordDetection K={K} (block contract, delayed output) -> "
                               + ("" if args.no_cfc else "CoarseFrequencyCorrection(delay 26) fused into ") +
                               f"SymbolFilter 32x44 -> SyncwordWipeoff fused into CostasLoop (BPSK, B_L T 0.01) over a "
                               f"2^{log2n}-sample synthetic cf32 capture on 1 B200, Es/N0 {args.esn0:g} dB",
                   "samples_per_gpu": n, "l2": "streams (8 B/sample) far larger than L2; no flush needed"},
        "detections_per_step": ndet, "symbols_per_step": n_sym, "symbol_tags_per_step": ntags,
        "gpu_launches": int(launches),
        "stage_ms": {"frontend": fe_ms, "syncword_detection": sd_ms, "symbol_filter_incl_host_plan": sf_ms,
                     "wipeoff_costas_loop": cl_ms},
        "frontend_fp_contract": {"ms": fe_fma_ms, "value_with_it": c_in / ((ms_per_step - fe_ms + fe_fma_ms) * 1e-3) / 1e6,
                                 "note": "opt-in b200sync_fe_config::fp_contract = 1: fused multiply-add per tap (one rounding "
                                         "instead of the reference's two, relative L2 difference ~1e-7); `value` above is the "
                                         "bit-exact default, value_with_it = the same step with this front end instead"},
        "roofline": hbm_roofline("frontend_kernel", 16.0 * n_out, fe_ms, peaks,
                                 "16 B/sample (8 in + 8 out); 2 x 40 taps x (rounded packed product, packed add) = 160 "
                                 "FP32x2 instructions per output (bit-exact std::inner_product order) + ~150 of timing, "
                                 "rotator and staging: FP32 pipe 52 % / issue 70 %",
                                 bound="fp32_issue"),
        "roofline_syncword_detection": hbm_roofline("correlate_kernel + peak stage + refine", 16.0 * consumed, sd_ms,
                                                    peaks, binding_roof(K)[1], bound=binding_roof(K)[0]),
        "roofline_symbol_filter": hbm_roofline("symbol_filter_kernel", 10.0 * consumed, sf_ms, peaks,
                                               "10 B/sample (8 in + 8/4 out); stage time includes the host replay of "
                                               "the tag state machine and the segment upload", bound="latency+hbm"),
        "roofline_costas_loop": hbm_roofline("costas_kernel", 16.0 * n_sym, cl_ms, peaks,
                                             "16 B/symbol (8 in + 8 out), one thread per packet stretch: bound by the "
                                             "latency of the longest sequential recurrence (one packet), not by HBM",
                                             bound="latency"),
    }
    rec["e2e"] = None
    if e2e:
        # raw capture in pinned host memory -> H2D -> the chain -> symbols back to pinned host memory
        hraw = torch.empty(n, dtype=torch.complex64, pin_memory=True)
        hraw.copy_(raw)
        hsym = torch.empty(sym.numel(), dtype=torch.complex64, pin_memory=True)
        torch.cuda.synchronize()

        def step_host():
            raw.copy_(hraw, non_blocking=True)
            r = step()
            hsym[:r[4]].copy_(sym[:r[4]], non_blocking=True)
            torch.cuda.synchronize()
            return r

        step_host()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            r = step_host()
        sec = (time.perf_counter() - t0) / reps
        rec["e2e"] = {"value": r[0] / sec / 1e6, "unit": "Msps", "h2d_bytes_per_step": int(n * 8),
                      "d2h_bytes_per_step": int(r[4] * 8 + r[3] * 48),
                      "note": "copies not overlapped with the chain (whole capture up, chain, symbols down)"}
        del hraw, hsym
    rec["cpu_baseline"] = None
    if cpu:
        th = os.cpu_count() or 1
        v = cpu_chain_rate(1 << 20, th, args.bins, args.esn0, args.thr)
        if v is not None:
            rec["cpu_baseline"] = {"value": v, "unit": "Msps", "cores": th, "kind": "reference",
                                   "sample": f"{th} independent chains x 2^20 raw samples, the reference's own classes "
                                             "(oracle/_ref/librefblocks.so) wired as apps/packet_transceiver.cpp + "
                                             "PM/packet_receiver.hpp do; radix-2 FFT stand-in for FFTW"}
    del raw, y, dl, sym
    torch.cuda.empty_cache()
    return rec


def run_chain(args):
    import torch

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("--workload chain is a single-GPU workload")
    torch.cuda.set_device(0)
    W = max(args.warmup, 3)
    sampler = ClockSampler(0)
    sampler.start()
    sampler.mark_begin()
    rec = measure_chain(args, args.log2n, args.steps, W, e2e=not args.no_e2e, cpu=not args.no_cpu)
    clocks = sampler.stop()
    line = {"metric": rec.pop("metric"), "value": rec.pop("value"), "unit": rec.pop("unit"), "n_gpus": 1,
            "steps": args.steps, "warmup": W, "ms_per_step": rec.pop("ms_per_step"), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "clocks": clocks}
    rec.pop("steps", None)
    line.update(rec)
    print(json.dumps(line))


def run_channels(args):
    """BASELINE configs[4], channel mode: `--channels` independent channels x 2^log2n samples.  Under torchrun
    the channels are partitioned over the ranks (channel c on rank c mod world; strong scaling: the total work
    is fixed), each rank makes one batched call per step; nothing is exchanged on the data path, the records
    stay with the rank that owns the channel."""
    import torch
    import torch.distributed as dist

    from gr4_packet_modem_b200 import SyncwordDetection, _native
    from gr4_packet_modem_b200.stimulus import DeviceStimulus

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    C_, n = args.channels, 1 << args.log2n
    mine = list(range(rank, C_, world))   # this rank's channels
    Cl = len(mine)
    K = 2 * args.bins + 1
    W = max(args.warmup, 3)
    lib = _native.lib()
    stream = torch.cuda.current_stream().cuda_stream
    x = torch.empty(max(Cl, 1) * n, dtype=torch.complex64, device=dev)
    for i, c in enumerate(mine):
        DeviceStimulus(seed=100 + c, esn0_db=args.esn0, cfo=0.005, device=local).generate_device(
            x[i * n:].data_ptr(), n, 0, stream)
    torch.cuda.synchronize()
    sd = SyncwordDetection(**rx_settings(args.bins, args.thr), device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if Cl == 0:
            return 0, []
        return sd.detect_channels_device(x.data_ptr(), Cl, n, n, stream)

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(W):
        consumed, per = step()
    barrier()
    launches0 = lib.b200sync_launch_count()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        consumed, per = step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = lib.b200sync_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    total = Cl * consumed
    ndet = int(sum(len(p) for p in per))
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        c = torch.tensor([total, ndet], device=dev, dtype=torch.int64)
        dist.all_reduce(c)
        total, ndet = int(c[0].item()), int(c[1].item())
    ms_per_step = ms / args.steps
    # single-stream runs of the same channels, one after the other, for comparison
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    same = True
    for i in range(Cl):
        _, r, _ = sd.detect_device(x.data_ptr() + 8 * i * n, n, stream)
        same = same and np.array_equal(r.view(np.uint8), per[i].view(np.uint8))
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        f = torch.tensor([1 if same else 0], device=dev, dtype=torch.int64)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        same = bool(f.item())
    if rank != 0:
        dist.destroy_process_group()
        return
    peaks = load_peaks()
    line = {
        "metric": "complex Msps (cf32) through RX sync (SyncwordDetection, batched channels)",
        "value": total / (ms_per_step * 1e-3) / 1e6, "unit": "Msps", "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[4] channel mode: {C_} independent channels x 2^{args.log2n} samples, "
                               f"K={K}, channels partitioned over {world} B200(s), one "
                               f"b200sync_sd_detect_channels_device call per rank and step",
                   "channels": C_, "samples_per_channel": n,
                   "l2": "capture (8 B/sample x channels) far larger than L2; no flush needed"},
        "detections_per_step": ndet, "clocks": clocks, "gpu_launches": int(launches),
        "e2e": None,
        "sequential_single_stream_ms_rank0": t0.elapsed_time(t1), "batched_equals_single_stream": bool(same),
        "roofline": hbm_roofline("correlate_kernel (whole step, rank 0)", 8.0 * Cl * consumed, ms_per_step, peaks,
                                 "whole-step figure; K>=3 is FP32/shared-memory bound (see the detect workload)"),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_detect(args, bins: int, log2n: int, esn0: float, cfo: float, steps: int, warmup: int, with_output: bool,
                   dev_index: int = 0) -> dict:
    """One compact single-GPU detection record (a `configs` entry of the N = 1 line): device-resident capture,
    stage times from the library's CUDA events."""
    import torch

    from gr4_packet_modem_b200 import SyncwordDetection
    from gr4_packet_modem_b200.stimulus import DeviceStimulus

    dev = torch.device("cuda", dev_index)
    n = 1 << log2n
    K = 2 * bins + 1
    stream = torch.cuda.current_stream().cuda_stream
    x = DeviceStimulus(seed=1, esn0_db=esn0, cfo=cfo, device=dev_index).generate(n, dev)
    out = torch.empty(n, dtype=torch.complex64, device=dev) if with_output else None
    sd = SyncwordDetection(**rx_settings(bins, args.thr), device=dev_index)
    tm = []
    for i in range(warmup + steps):
        if i == warmup:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        c, recs, _ = sd.detect_device(x.data_ptr(), n, stream, d_out_ptr=out.data_ptr() if with_output else 0, copy=False)
        if i >= warmup:
            tm.append(sd.last_timings())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    cm = statistics.mean(t["correlate_ms"] for t in tm)
    bps = 16 if with_output else 8
    bound, note = binding_roof(K)
    rec = {"value": c / (ms * 1e-3) / 1e6, "unit": "Msps", "ms_per_step": ms, "steps": steps, "K": K,
           "samples": n, "esn0_db": esn0, "cfo": cfo, "bytes_per_sample": bps, "detections_per_step": len(recs),
           "stage_ms": {"correlate": cm, "peaks": statistics.mean(t["peaks_ms"] for t in tm),
                        "refine_and_copy": statistics.mean(t["refine_ms"] for t in tm)},
           "roofline": hbm_roofline("correlate_kernel", float(bps) * c, cm, load_peaks(), note, bound=bound),
           "fp32_tflops_nominal": c * flop_per_sample(K) / (cm * 1e-3) / 1e12}
    rec["roofline"]["traffic"] = measured_traffic(log2n, K, "block" if with_output else "detect")
    rec["roofline_whole_step"] = hbm_roofline("whole step", float(bps) * c, ms, load_peaks(),
                                              "all kernels of the step + record copy", bound=bound)
    del x, out, sd
    torch.cuda.empty_cache()
    return rec


def measure_channels(args, channels: int, log2n: int, steps: int, warmup: int) -> dict:
    import torch

    from gr4_packet_modem_b200 import SyncwordDetection, _native
    from gr4_packet_modem_b200.stimulus import DeviceStimulus

    dev = torch.device("cuda", 0)
    n = 1 << log2n
    stream = torch.cuda.current_stream().cuda_stream
    lib = _native.lib()
    x = torch.empty(channels * n, dtype=torch.complex64, device=dev)
    for c in range(channels):
        DeviceStimulus(seed=100 + c, esn0_db=args.esn0, cfo=0.005).generate_device(x[c * n:].data_ptr(), n, 0, stream)
    torch.cuda.synchronize()
    sd = SyncwordDetection(**rx_settings(args.bins, args.thr), device=0)
    for _ in range(warmup):
        consumed, per = sd.detect_channels_device(x.data_ptr(), channels, n, n, stream)
    torch.cuda.synchronize()
    l0 = lib.b200sync_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        consumed, per = sd.detect_channels_device(x.data_ptr(), channels, n, n, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    K = 2 * args.bins + 1
    bound, note = binding_roof(K)
    rec = {"value": channels * consumed / (ms * 1e-3) / 1e6, "unit": "Msps", "ms_per_step": ms, "steps": steps,
           "channels": channels, "samples_per_channel": n, "K": K,
           "detections_per_step": int(sum(len(p) for p in per)),
           "gpu_launches_per_step": int((lib.b200sync_launch_count() - l0) // steps),
           "roofline": hbm_roofline("whole step (all channels)", 8.0 * channels * consumed, ms, load_peaks(), note,
                                    bound=bound)}
    del x, sd
    torch.cuda.empty_cache()
    return rec


def bind_to_gpu_cores(local: int) -> str:
    """Pin this rank (and the pinned buffers it allocates afterwards: first touch) to the host cores local to its
    GPU — /sys/bus/pci/devices/<bus id>/local_cpulist.  Best effort; returns what was done for the JSON line."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for tok in spec.split(","):
            a, _, b = tok.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"{bus}: no allowed core in local_cpulist {spec}"
        os.sched_setaffinity(0, cpus)
        node = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
        return f"{bus}: numa node {node}, cores {spec}"
    except Exception as e:
        return f"not bound ({e!r})"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=30, help="log2 samples per GPU")
    ap.add_argument("--bins", type=int, default=4, help="min/max_freq_bin = -/+bins (K = 2*bins+1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the compact records of the other BASELINE configs")
    ap.add_argument("--no-output", action="store_true", help="detection only: do not write the delayed output span (8 B/sample)")
    ap.add_argument("--no-bind", action="store_true", help="do not bind ranks to their GPU's local host cores")
    ap.add_argument("--workload", default="detect", choices=["detect", "chain", "channels"],
                    help="detect: BASELINE configs[1] (default, the metric's configuration; --bins 16 --esn0 0 gives "
                         "configs[3]); chain: configs[2]; channels: configs[4] channel mode")
    ap.add_argument("--esn0", type=float, default=20.0, help="Es/N0 of the synthetic capture in dB")
    ap.add_argument("--thr", type=float, default=9.5, help="power_threshold")
    ap.add_argument("--cfo", type=float, default=0.005, help="carrier frequency offset of the capture, rad/sample")
    ap.add_argument("--channels", type=int, default=64)
    ap.add_argument("--no-cfc", action="store_true", help="chain workload: leave CoarseFrequencyCorrection out")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "chain":
        return run_chain(args)
    if args.workload == "channels":
        if args.log2n == 30:
            args.log2n = 24
        return run_channels(args)

    import torch
    import torch.distributed as dist

    from gr4_packet_modem_b200 import SyncwordDetection, _native
    from gr4_packet_modem_b200.stimulus import DeviceStimulus

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    binding = "off (--no-bind)" if args.no_bind else bind_to_gpu_cores(local)
    if world > 1:
        # CUDA tensors (the contract's barrier) over NCCL, CPU tensors (chain tables, timings) over gloo: the
        # data path itself has no collective at all
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = 2 * args.bins + 1
    n_per = 1 << args.log2n
    S = FFT - 297 + 1
    with_out = not args.no_output
    bps = 16 if with_out else 8
    sd = SyncwordDetection(**rx_settings(args.bins, args.thr), device=local)
    assert sd.stride == S
    stream = torch.cuda.current_stream().cuda_stream
    lib = _native.lib()

    # ---- this rank's time shard of a world*n_per-sample capture (shard + halo blocks) ----
    from gr4_packet_modem_b200.sharding import SharedTableExchange, gather_entry_offset, plan_shards

    # the shards' chain tables travel as host bytes: through a shared-memory segment when all ranks share the box
    # (they do: one node), else over the gloo backend.  Either way nothing touches a GPU or NCCL.
    xchg = None
    if world > 1:
        try:
            xchg = SharedTableExchange(f"b200sync_{os.environ.get('MASTER_PORT', '0')}_{os.getuid()}", rank, world, TAU + 1)
        except Exception as e:
            sys.stderr.write(f"rank {rank}: shared-memory exchange unavailable ({e!r}); using gloo\n")
        ok = torch.tensor([1 if xchg is not None else 0], dtype=torch.int64)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            xchg = None

    def entry_offset(table):
        return xchg.entry_offset(table) if xchg is not None else gather_entry_offset(table, rank, world)

    total_n = n_per * world
    shard = plan_shards(total_n, world, FFT, S, TAU)[rank]
    total_blocks, fb, nbk = shard.total_blocks, shard.first_block, shard.n_blocks
    seg0 = shard.first_sample
    seg_n = shard.n_samples if world > 1 else n_per
    # the capture is written into HBM by the native generator (csrc/stimulus.cu): every sample is a function
    # of (seed, absolute index), so each rank generates exactly its own shard + halo of the same stream
    x = DeviceStimulus(seed=1, esn0_db=args.esn0, cfo=args.cfo, device=local).generate(seg_n, dev, seg0)
    # this rank's slice of the block's output span: output items [fb*S, (fb+nbk)*S)
    out_first, out_len = fb * S, nbk * S
    d_out = torch.empty(max(out_len, 1), dtype=torch.complex64, device=dev) if with_out else None
    torch.cuda.synchronize()
    max_recs = seg_n // (TAU + 1) + 2

    def step_device():
        if world == 1:
            # (records and tags as views into the context's buffers — valid until the next call, like processBulk's spans)
            c, recs, _ = sd.detect_device(x.data_ptr(), n_per, stream, d_out_ptr=d_out.data_ptr() if with_out else 0, copy=False)
            return c, len(recs)
        if with_out:
            sd.shard_output(d_out.data_ptr(), out_first, out_len)
        table = sd.shard_phase1(x.data_ptr(), seg0, seg_n, fb, nbk, total_blocks, stream)
        # T+1 small integers per rank, exchanged as host bytes: the only exchange of the path
        j = entry_offset(table)
        recs, _ = sd.shard_phase2(j, max_recs, copy=False)
        return nbk * S, len(recs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(*vals):
        if world == 1:
            return vals
        t = torch.tensor(list(vals), dtype=torch.int64)
        dist.all_reduce(t)
        return tuple(int(v) for v in t.tolist())

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(W):
        step_device()
    barrier()
    launches0 = lib.b200sync_launch_count()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    consumed = ndet = 0
    corr_ms = []
    for _ in range(args.steps):
        consumed, ndet = step_device()
        corr_ms.append(sd.last_timings())
    shard_samples = consumed  # this rank's samples per step: what ITS correlate launch processed
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = lib.b200sync_launch_count() - launches0
    ms = max_over_ranks(e0.elapsed_time(e1))
    consumed, ndet = sum_over_ranks(consumed, ndet)
    ms_per_step = ms / args.steps
    value = consumed / (ms_per_step * 1e-3) / 1e6  # Msps, whole job

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region: the capture
    # goes up over PCIe, records AND the block's delayed output span come back in host memory
    e2e = None
    if not args.no_e2e:
        hx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
        hx.copy_(x)
        hout = torch.empty(max(out_len if world > 1 else n_per, 1), dtype=x.dtype, pin_memory=True) if with_out else None
        torch.cuda.synchronize()
        n_host = n_per if world == 1 else seg_n

        def step_host():
            if world == 1:
                c, recs, _ = sd.detect_host((hx.data_ptr(), n_host), out=hout.data_ptr() if with_out else None)
                return c, len(recs)
            # this rank's shard + halo from pinned host memory: the correlator chases the H2D pieces
            if with_out:
                sd.shard_output_host(hout.data_ptr(), out_first, out_len)
            table = sd.shard_phase1_host((hx.data_ptr(), n_host), seg0, fb, nbk, total_blocks)
            j = entry_offset(table)
            recs, _ = sd.shard_phase2(j, max_recs)
            return nbk * S, len(recs)

        def timed_host(reps=3):
            step_host()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                step_host()
            barrier()
            return max_over_ranks((time.perf_counter() - t0) / reps)

        sec = timed_host()
        e2e = {"value": consumed / sec / 1e6, "unit": "Msps", "h2d_bytes_per_step": int(n_host * 8 * world),
               "d2h_bytes_per_step": int(ndet * 48 + 16 * world + (consumed * 8 if with_out else 0)),
               "h2d_gbs_per_gpu": n_host * 8 / sec / 1e9, "host_binding_rank0": binding,
               "note": "block contract with HOST spans: input from pinned host memory over PCIe (the correlator chases "
                       "the copies), the delayed output span written on the device and copied back to pinned host "
                       "memory on a second stream (PCIe is full duplex), records back over PCIe"}
        if with_out:   # the same without the output span: what round 1 reported as e2e
            with_out_saved, with_out = with_out, False
            sec2 = timed_host(2)
            with_out = with_out_saved
            e2e["detection_only"] = {"value": consumed / sec2 / 1e6, "unit": "Msps", "h2d_gbs_per_gpu": n_host * 8 / sec2 / 1e9,
                                     "note": "no output span: 8 B/sample up, records down"}
        del hx, hout

    if xchg is not None:
        barrier()
        xchg.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = None
    extra = {}
    if corr_ms:
        cm = statistics.mean(d["correlate_ms"] for d in corr_ms)
        pm = statistics.mean(d["peaks_ms"] for d in corr_ms)
        rm = statistics.mean(d["refine_ms"] for d in corr_ms)
        bound, note = binding_roof(K)
        ach = shard_samples * bps / (cm * 1e-3) / 1e9  # rank 0's launch (every rank runs the same shape)
        roofline = {"bound": bound, "kernel": "correlate_kernel", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": measured_traffic(args.log2n, K, "block" if with_out else "detect"),
                    "algorithmic_bytes_per_launch": shard_samples * bps, "ms_per_launch": cm,
                    "peak_source": "MEASURED_PEAKS.json (burst copy)" if peaks else "fallback 6650 GB/s",
                    "note": "achieved/peak/frac are the HBM figures the contract asks for (algorithmic "
                            f"{bps} B/sample x samples of the launch / its device time); `bound` names what binds: " + note}
        fp32_peak = measured_fp32_peak()
        ach_tf = shard_samples * flop_per_sample(K) / (cm * 1e-3) / 1e12
        extra = {"stage_ms": {"correlate": cm, "peaks": pm, "refine_and_copy": rm},
                 "fp32": {"achieved_tflops": ach_tf, "nominal_peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12,
                          "measured_peak_tflops": fp32_peak, "frac_of_measured": ach_tf / fp32_peak if fp32_peak else None,
                          "note": "achieved = nominal flop (5 N log2 N per FFT, SURVEY §8d) / correlate time; measured "
                                  "peak = FFMA2 microbenchmark, profiles/r1_fp32_peak.json (an FMA = 2 flop)"}}

    cpu = None
    if world == 1 and not args.no_cpu:
        th = os.cpu_count() or 1
        kind = cpu_reference_kind()
        r1, _ = cpu_reference_rate(args.bins, 1 << 22, 1, 1, 1, args.esn0, args.thr, kind)
        rN, _ = cpu_reference_rate(args.bins, 1 << 22, th, 1, 1, args.esn0, args.thr, kind)
        rP, _ = cpu_reference_rate(args.bins, 1 << 22, 1, 1, 1, args.esn0, args.thr, "port")
        cpu = {"value": rN, "unit": "Msps", "cores": th, "kind": kind, "single_core_msps": r1,
               "oracle_port_single_core_msps": rP,
               "sample": f"{th} independent streams x 2^22 samples of the same signal model; " + CPU_KIND_NOTE[kind]}

    # ---- the other BASELINE configurations, compactly, in the same driver-run line (N = 1 only)
    configs = None
    if world == 1 and not args.no_configs:
        del x, d_out
        torch.cuda.empty_cache()
        configs = {}
        sub = [("k1_detection_only", lambda: measure_detect(args, 0, args.log2n, 20.0, 0.0, 5, 3, False)),
               ("k1_block_contract", lambda: measure_detect(args, 0, args.log2n, 20.0, 0.0, 5, 3, True)),
               ("k9_detection_only", lambda: measure_detect(args, 4, args.log2n, 20.0, 0.005, 5, 3, False)),
               ("k17_0dB", lambda: measure_detect(args, 8, min(args.log2n, 28), 0.0, 0.005, 3, 3, True)),
               ("k33_0dB", lambda: measure_detect(args, 16, min(args.log2n, 28), 0.0, 0.005, 3, 3, True)),
               ("chain", lambda: measure_chain(args, min(args.log2n, 29), 3, 3, e2e=not args.no_e2e, cpu=not args.no_cpu)),
               ("channels64", lambda: measure_channels(args, 64, min(args.log2n - 6, 24), 3, 3))]
        for name, fn in sub:
            try:
                configs[name] = fn()
            except Exception as e:  # a sub-record must never cost the headline
                configs[name] = {"error": repr(e)}
                torch.cuda.empty_cache()
        configs["_what"] = ("k1_*: bins 0 (benchmarks/results.md:37), CFO 0 so the single hypothesis sees the packets; "
                            "k9_detection_only: the round-1 headline configuration (no output span, 8 B/sample); "
                            "k17/k33_0dB: BASELINE configs[3]; chain: configs[2]; channels64: configs[4] channel mode")

    line = {
        "metric": "complex Msps (cf32) through RX sync (SyncwordDetection)", "value": value, "unit": "Msps",
        "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.log2n, K, args.esn0, args.thr),
        "detections_per_step": ndet, "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
        "roofline": roofline, "cpu_baseline": cpu,
    }
    line.update(extra)
    if configs is not None:
        line["configs"] = configs
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
