"""py3 compat loader for the UNMODIFIED reference source tree (test infrastructure only).

This is TEST INFRASTRUCTURE.  Nothing in the product path (afterqc_b200/) may import it.
It only works where /root/reference exists (the build container); it never travels to the
GPU box.  It is used to (1) pin the C restatement in oracle/aqc_oracle.c against the real
reference and (2) generate the committed golden vectors under tests/golden/
(see oracle/make_golden.py).

The reference is Python 2 (after.py:189-191 refuses py3).  The loader compiles each reference
module from the source text where it lies (read-only), applying the in-memory shims listed in
SURVEY.md section 8(c):

  1. xrange -> range                       (util.py:44, preprocesser.py:38, qualitycontrol.py:43 ...)
  2. py3 guard in after.main bypassed      (after.py:189-191) -- we call processOptions directly
  3. gzip/bz2 opened in latin-1 text mode  (fastq.py:24,26,68 rely on py2 str I/O)
  4. py2 integer division sites -> //      (preprocesser.py:752, qualitycontrol.py:241,246)
  5. util.EDIT_DISTANCE_MODULE_EXISTS forced False (the editdistance/ dir is a spurious
     namespace package under py3; util.py:11-24)
  6. matplotlib absent/ignored             (qualitycontrol.py:11-21, dead plotting code)

No reference source is copied into this repository: modules are compiled from
REFERENCE_DIR at import time.
"""
import builtins
import bz2 as _bz2
import gzip as _gzip
import io
import os
import sys
import types

REFERENCE_DIR = os.environ.get("AFTERQC_REFERENCE_DIR", "/root/reference")

_MODULES = ["util", "fastq", "qualitycontrol", "qcreporter", "barcodeprocesser", "preprocesser", "after"]


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "preprocesser.py"))


class _GzipShim(types.ModuleType):
    """gzip whose open() defaults to py2-style str I/O."""

    def __init__(self):
        super().__init__("gzip")
        self.__dict__.update({k: v for k, v in _gzip.__dict__.items() if k != "open"})

    @staticmethod
    def open(filename, mode="r", compresslevel=9):
        if "b" not in mode and "t" not in mode:
            mode = mode + "t"
        if "t" in mode:
            return _gzip.open(filename, mode, compresslevel=compresslevel, encoding="latin-1", newline="\n")
        return _gzip.open(filename, mode, compresslevel=compresslevel)


class _Bz2Shim(types.ModuleType):
    def __init__(self):
        super().__init__("bz2")
        self.__dict__.update({k: v for k, v in _bz2.__dict__.items() if k != "BZ2File"})

    @staticmethod
    def BZ2File(filename, mode="r"):
        return _bz2.open(filename, mode + "t", encoding="latin-1", newline="\n")


_PATCHES = {
    "preprocesser": [("float(OVERLAP_LEN_SUM/OVERLAPPED)", "float(OVERLAP_LEN_SUM//OVERLAPPED)")],
    "qualitycontrol": [
        ("len(self.topKmerCount)/2)", "len(self.topKmerCount)//2)"),
        ("(len(self.topKmerCount) - shift) / top", "(len(self.topKmerCount) - shift) // top"),
    ],
    "barcodeprocesser": [],
    "util": [],
}

_loaded = None


def load():
    """Return a dict name -> module for the reference modules, compiled with the py3 shims."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s (ref_loader only works in the build container)" % REFERENCE_DIR)
    builtins.xrange = range  # shim 1
    mods = {}
    saved = {}
    shim_names = {"gzip": _GzipShim(), "bz2": _Bz2Shim()}
    # the reference imports siblings by bare name; expose them under those names while loading
    for name in _MODULES + list(shim_names):
        saved[name] = sys.modules.get(name)
    try:
        sys.modules["gzip"] = shim_names["gzip"]
        sys.modules["bz2"] = shim_names["bz2"]
        sys.modules["matplotlib"] = None  # shim 6: force the HAVE_MATPLOTLIB=False path
        for name in _MODULES:
            path = os.path.join(REFERENCE_DIR, name + ".py")
            with io.open(path, "r", encoding="latin-1") as f:
                src = f.read()
            for old, new in _PATCHES.get(name, []):
                assert old in src, "patch site vanished in %s: %r" % (name, old)
                src = src.replace(old, new)
            mod = types.ModuleType(name)
            mod.__file__ = path
            sys.modules[name] = mod
            code = compile(src, path, "exec")
            exec(code, mod.__dict__)
            mods[name] = mod
        mods["util"].EDIT_DISTANCE_MODULE_EXISTS = False  # shim 5
    finally:
        sys.modules.pop("matplotlib", None)
        for name, m in saved.items():
            if m is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = m
    _loaded = mods
    return mods


def default_options(**overrides):
    """The optparse defaults of after.py:14-93 post-processed as after.main does (after.py:195-201)."""
    mods = load()
    after = mods["after"]
    argv = sys.argv
    try:
        sys.argv = ["after.py"]
        (options, _args) = after.parseCommand()
    finally:
        sys.argv = argv
    for k, v in overrides.items():
        if not hasattr(options, k):
            raise KeyError(k)
        setattr(options, k, v)
    options.version = after.AFTERQC_VERSION
    options.trim_pair_same = after.parseBool(options.trim_pair_same) if isinstance(options.trim_pair_same, str) else options.trim_pair_same
    options.draw = after.parseBool(options.draw) if isinstance(options.draw, str) else options.draw
    options.store_overlap = after.parseBool(options.store_overlap) if isinstance(options.store_overlap, str) else options.store_overlap
    options.trim_front2 = options.trim_front
    options.trim_tail2 = options.trim_tail
    if options.read1_file is not None and options.barcode_flag in options.read1_file and after.parseBool(options.barcode):
        options.barcode = True
        options.trim_front = 0
        options.trim_front2 = 0
    else:
        options.barcode = False
    return options


def run_cli(argv):
    """Run the reference CLI (after.main minus the py3 guard) on argv (list without program name)."""
    mods = load()
    after = mods["after"]
    saved = sys.argv
    try:
        sys.argv = ["after.py"] + list(argv)
        (options, _args) = after.parseCommand()
    finally:
        sys.argv = saved
    options.version = after.AFTERQC_VERSION
    options.trim_pair_same = after.parseBool(options.trim_pair_same)
    options.draw = after.parseBool(options.draw)
    options.store_overlap = after.parseBool(options.store_overlap)
    options.trim_front2 = options.trim_front
    options.trim_tail2 = options.trim_tail
    if options.input_dir is None and options.read1_file is None:
        options.input_dir = "."
    if options.input_dir is not None:
        after.processDir(options.input_dir, options)
    else:
        if options.barcode_flag in options.read1_file and after.parseBool(options.barcode):
            options.barcode = True
            options.trim_front = 0
            options.trim_front2 = 0
        else:
            options.barcode = False
        after.processOptions(options)
    return options
