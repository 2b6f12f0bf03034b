"""Kernel tuning harness: many build variants, ONE GPU call.

  python tools/variant_sweep.py build base= u4=-DAQC_WARPS=4 rare=-DAQC_NOINLINE_RARE ...     (here, no GPU: nvcc only)
  gpurun -- 'python tools/variant_sweep.py run [--pairs N] [--steps K] [--parity]'             (on the B200 box)

`build` compiles libafterqc_b200 once per variant into gpurun_variants/<name>.so (git-ignored, travels with gpurun) and
records the ptxas register/spill lines; `run` points AQC_LIB_PATH at each variant, optionally runs the filter/stat parity
tests against the oracle first, then `bench.py --no-e2e --no-cpu`, and writes gpurun_out/variant_sweep.json plus a
table (M read-pairs/s, filter-kernel ms).  A variant that fails parity is reported and not benchmarked."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "gpurun_variants")


def build(specs):
    from afterqc_b200 import build as B
    os.makedirs(VDIR, exist_ok=True)
    meta = {}
    for spec in specs:
        name, _, flags = spec.partition("=")
        out = os.path.join(VDIR, name + ".so")
        cmd = [B.nvcc_path()] + B.NVCC_FLAGS + ["-Xptxas", "-v"] + [f for f in flags.split(",") if f] + ["-o", out] + B.SOURCES + ["-lz", "-lpthread"]
        p = subprocess.run(cmd, cwd=B.CSRC, capture_output=True, text=True)
        if p.returncode:
            print("variant %s: BUILD FAILED\n%s" % (name, p.stderr[-2000:]))
            if os.path.exists(out):
                os.unlink(out)
            continue
        regs = re.findall(r"Used (\d+) registers", p.stderr)
        spills = re.findall(r"(\d+) bytes spill stores", p.stderr)
        meta[name] = {"flags": flags, "max_registers": max(map(int, regs)) if regs else None,
                      "max_spill_store_bytes": max(map(int, spills)) if spills else None}
        print("variant %-16s flags=%-40s regs(max)=%s spill(max)=%s" % (name, flags, meta[name]["max_registers"], meta[name]["max_spill_store_bytes"]))
    with open(os.path.join(VDIR, "variants.json"), "w") as f:
        json.dump(meta, f, indent=1)


def run(argv):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=10000000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--parity", action="store_true")
    ap.add_argument("--filter-kernel", default="warp", choices=["warp", "lane"],
                    help="lane: benchmark lane_kernel (variants: -DAQC_LANE_TWO_PLANE_FILTER, -DAQC_LANE_UNROLL_CONVERT; env AQC_LANE_WARPS); "
                         "--parity then runs tests/lane_gpu_check.py parity instead of the pair_kernel tests")
    a = ap.parse_args(argv)
    with open(os.path.join(VDIR, "variants.json")) as f:
        meta = json.load(f)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    rows = []
    for name in meta:
        env = dict(os.environ, AQC_LIB_PATH=os.path.join(VDIR, name + ".so"))
        row = {"variant": name, **meta[name]}
        if a.parity:
            if a.filter_kernel == "lane":
                p = subprocess.run([sys.executable, "tests/lane_gpu_check.py", "parity"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
            else:
                p = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-x", "-q", "-k", "filter_parity or stat_parity or single_end"],
                                   cwd=ROOT, env=env, capture_output=True, text=True)
            row["parity"] = "pass" if p.returncode == 0 else "FAIL"
            if p.returncode:
                row["parity_tail"] = p.stdout[-400:]
                rows.append(row)
                continue
        p = subprocess.run([sys.executable, "bench.py", "--pairs", str(a.pairs), "--steps", str(a.steps), "--warmup", "3", "--no-e2e", "--no-cpu",
                            "--filter-kernel", a.filter_kernel],
                           cwd=ROOT, env=env, capture_output=True, text=True)
        try:
            j = json.loads(p.stdout.strip().splitlines()[-1])
            row.update(value=j["value"], ms_per_step=j["ms_per_step"], kernel_ms=j["roofline"]["kernel_ms"], sm_mhz=j["clocks"]["sm_mhz"])
        except Exception as e:
            row["bench_error"] = "%s: %s" % (e, p.stderr[-300:])
        rows.append(row)
        print(row, flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "variant_sweep.json"), "w") as f:
        json.dump(rows, f, indent=1)
    print("\n%-16s %-8s %10s %10s %6s %6s" % ("variant", "parity", "Mpairs/s", "kernel ms", "regs", "spill"))
    for r in rows:
        print("%-16s %-8s %10s %10s %6s %6s" % (r["variant"], r.get("parity", "-"), ("%.1f" % r["value"]) if "value" in r else "-",
                                                  ("%.3f" % r["kernel_ms"]) if "kernel_ms" in r else "-", r.get("max_registers"), r.get("max_spill_store_bytes")))


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "build":
        build(sys.argv[2:])
    elif len(sys.argv) >= 2 and sys.argv[1] == "run":
        run(sys.argv[2:])
    else:
        print(__doc__)
