"""SASS fingerprints of the kernels in libafterqc_b200.so.

The kernels that have been verified on a B200 must not change behind a refactoring of shared headers: this tool hashes the
instruction stream of every kernel (`cuobjdump -sass`, addresses and encodings included) and compares it with
profiles/sass_fingerprints.json, which lists the kernels whose parity was seen green ON HARDWARE together with the nvcc
version that built them.

  python tools/sass_fingerprint.py            compare (exit 1 on a difference in a listed kernel)
  python tools/sass_fingerprint.py --update   rewrite the file for the listed kernels (after re-verifying them on the GPU)
  python tools/sass_fingerprint.py --relist   list EVERY kernel of the library (after the whole GPU suite passed on this build)
  python tools/sass_fingerprint.py --all      print every kernel's size and hash
"""
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "afterqc_b200", "libafterqc_b200.so")
FILE = os.path.join(ROOT, "profiles", "sass_fingerprints.json")


def nvcc_version():
    try:
        out = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout
        m = re.search(r"release [\d.]+, V([\d.]+)", out)
        return m.group(1) if m else "unknown"
    except Exception:       # noqa: BLE001
        return "unknown"


def fingerprints(lib=LIB):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
            funcs[cur].append(ln.strip())
    return {k: {"instructions": len(v), "sha1": hashlib.sha1("\n".join(v).encode()).hexdigest()} for k, v in funcs.items()}


def demangled(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:       # noqa: BLE001
        return name


def compare():
    """-> list of problems (empty = the listed kernels are unchanged)"""
    with open(FILE) as f:
        want = json.load(f)
    if want.get("nvcc") != nvcc_version():
        return []           # another compiler: the hashes say nothing
    got = fingerprints()
    by_hash = {v["sha1"]: k for k, v in got.items()}
    problems = []
    for name, w in want["kernels"].items():
        g = got.get(name)
        if g is not None and g["sha1"] == w["sha1"]:
            continue
        if w["sha1"] in by_hash:        # same instruction stream under another name (a template parameter was added or retyped)
            continue
        if g is None:
            problems.append("%s: neither the name nor its instruction stream is in the library any more" % w.get("name", name))
        else:
            problems.append("%s: SASS changed (%d -> %d instructions)" % (w.get("name", name), w["instructions"], g["instructions"]))
    return problems


if __name__ == "__main__":
    if "--all" in sys.argv:
        for k, v in sorted(fingerprints().items()):
            print("%6d %s %s" % (v["instructions"], v["sha1"][:12], demangled(k)))
    elif "--relist" in sys.argv:
        got = fingerprints()
        note = "kernels of libafterqc_b200.so whose parity tests (pytest -m gpu) were green on a B200 with exactly this instruction stream"
        want = {"note": note, "nvcc": nvcc_version(),
                "kernels": {k: dict(v, name=demangled(k)) for k, v in sorted(got.items()) if "aqc" in k}}
        with open(FILE, "w") as f:
            json.dump(want, f, indent=1, sort_keys=True)
        print("relisted %d kernels in %s" % (len(want["kernels"]), FILE))
    elif "--update" in sys.argv:
        with open(FILE) as f:
            want = json.load(f)
        got = fingerprints()
        for name in list(want["kernels"]):
            if name in got:
                want["kernels"][name].update(got[name])
        want["nvcc"] = nvcc_version()
        with open(FILE, "w") as f:
            json.dump(want, f, indent=1, sort_keys=True)
        print("updated", FILE)
    else:
        p = compare()
        print("\n".join(p) if p else "listed kernels unchanged")
        sys.exit(1 if p else 0)
