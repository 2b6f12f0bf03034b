"""Command line of the B200 engine: flag-for-flag compatible with the reference's after.py
(after.py:14-93 options, :186-225 main, :101-171 directory mode) so that it is a drop-in for
the per-read path.  Options are declared as a table; semantics (defaults, string booleans via
parseBool util.py:29-34, trim_front2/trim_tail2 mirroring, barcode detection by file name) follow
the reference.  Index files and barcoded (UMI) files are handled on the host around the device loop; debubble is
outside the scope: accepted on the command line and refused at run time with a clear message.
"""
import copy
import os
import sys
import time
from multiprocessing import Process
from optparse import OptionParser

AFTERQC_VERSION = "0.9.6"

USAGE = ("Automatic Filtering, Trimming, Error Removing and Quality Control for Illumina fastq data "
         "(B200-native engine)\n\nSimplest usage:\ncd to the folder containing your fastq data, run <python after.py>")

# (short, long, kwargs)
_OPTIONS = [
    ("-1", "--read1_file", dict(help="file name of read1, required. If input_dir is specified, then this arg is ignored.")),
    ("-2", "--read2_file", dict(default=None, help="file name of read2, if paired.")),
    ("-7", "--index1_file", dict(default=None, help="file name of 7' index.")),
    ("-5", "--index2_file", dict(default=None, help="file name of 5' index.")),
    ("-d", "--input_dir", dict(default=None, help="the input dir to process automatically.")),
    ("-g", "--good_output_folder", dict(default="good", help="the folder to store good reads")),
    ("-b", "--bad_output_folder", dict(default=None, help="the folder to store bad reads")),
    ("-r", "--report_output_folder", dict(default=None, help="the folder to store QC reports")),
    ("", "--read1_flag", dict(default="R1", help="name flag of read1")),
    ("", "--read2_flag", dict(default="R2", help="name flag of read2")),
    ("", "--index1_flag", dict(default="I1", help="name flag of index1")),
    ("", "--index2_flag", dict(default="I2", help="name flag of index2")),
    ("-f", "--trim_front", dict(default=-1, type="int", help="bases trimmed at the head, -1 = auto")),
    ("-t", "--trim_tail", dict(default=-1, type="int", help="bases trimmed at the tail, -1 = auto")),
    ("", "--trim_pair_same", dict(default="true", help="same trimming for read1 and read2, default true")),
    ("-q", "--qualified_quality_phred", dict(default=15, type="int", help="qualified phred, default 15")),
    ("-u", "--unqualified_base_limit", dict(default=60, type="int", help="unqualified base limit, default 60")),
    ("-p", "--poly_size_limit", dict(default=35, type="int", help="polyX length limit, default 35")),
    ("-a", "--allow_mismatch_in_poly", dict(default=2, type="int", help="mismatches allowed in polyX, default 2")),
    ("-n", "--n_base_limit", dict(default=5, type="int", help="N base limit, default 5")),
    ("-s", "--seq_len_req", dict(default=35, type="int", help="minimum length, default 35")),
    ("", "--debubble", dict(action="store_true", default=False, help="debubble (out of scope for the B200 engine)")),
    ("", "--debubble_dir", dict(default="debubble", help="debubble output folder")),
    ("", "--draw", dict(default="on", help="draw pictures")),
    ("", "--barcode", dict(default="on", help="barcode handling for files with barcode_flag in the name")),
    ("", "--barcode_length", dict(default=12, type="int", help="barcode length")),
    ("", "--barcode_flag", dict(default="barcode", help="name flag of a barcoded file")),
    ("", "--barcode_verify", dict(default="CAGTA", help="barcode verify sequence")),
    ("", "--store_overlap", dict(default="off", help="store only overlapped bases of the good reads")),
    ("", "--overlap_output_folder", dict(default=None, help="folder for the overlapped bases")),
    ("", "--qc_only", dict(action="store_true", default=False, help="only QC result will be output")),
    ("", "--qc_sample", dict(default=200000, type="int", help="reads sampled for QC, 0 = all, default 200000")),
    ("", "--qc_kmer", dict(default=8, type="int", help="k-mer length for QC, default 8")),
    ("", "--no_correction", dict(action="store_true", default=False, help="disable base correction")),
    ("", "--mask_mismatch", dict(action="store_true", default=False, help="zero the quality of uncorrected mismatches")),
    ("", "--no_overlap", dict(action="store_true", default=False, help="disable overlap analysis")),
    ("-z", "--gzip", dict(action="store_true", default=False, help="force gzip output")),
    ("", "--compression", dict(type="int", default=2, help="gzip level 0-9, default 2")),
    # engine-only extras (not in the reference)
    ("", "--gpus", dict(type="int", default=1, help="[B200 engine] shard the reads over this many GPUs of the box")),
]


def parseBool(s):
    """util.py:29-34"""
    return s.lower() in ("true", "yes", "on")


def build_parser():
    parser = OptionParser(usage=USAGE, version=AFTERQC_VERSION)
    for short, long_, kw in _OPTIONS:
        parser.add_option(short, long_, dest=long_[2:], **kw)
    return parser


def parseCommand(argv=None):
    return build_parser().parse_args(argv)


def normalize_options(options):
    """after.py:194-201"""
    options.version = AFTERQC_VERSION
    for k in ("trim_pair_same", "draw", "store_overlap"):
        v = getattr(options, k)
        if isinstance(v, str):
            setattr(options, k, parseBool(v))
    options.trim_front2 = options.trim_front
    options.trim_tail2 = options.trim_tail
    return options


def matchFlag(filename, flag):
    """after.py:95-99"""
    if flag.endswith(('.', '_', '-')):
        return flag in filename
    return any((flag + sep) in filename for sep in "._-")


def processOptions(options):
    from .pipeline import seqFilter
    gpus = getattr(options, "gpus", 1) or 1
    if gpus > 1:
        from .multigpu import run_sharded
        return run_sharded(options, gpus)
    device = getattr(options, "device", -1)
    if device is None or device < 0:
        return seqFilter(options).run()
    from .engine import Engine
    seqFilter(options, backend_factory=lambda p: Engine(p, device=device)).run()


def device_count():
    """CUDA devices visible to the jobs (0 when the library or the driver is missing).  Asked in a child process: the jobs are
    forked (after.py:168-171) and CUDA must not have been initialised in the parent before a fork."""
    import subprocess
    code = "from afterqc_b200 import _native; print(_native.lib().aqc_device_count())"
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        return max(0, int(r.stdout.strip().splitlines()[-1]))
    except Exception:       # noqa: BLE001
        return 0


def processDir(folder, options):
    """after.py:101-171: one job per R1 file of the folder."""
    fqext = (".fq", ".fastq", ".fq.gz", ".fastq.gz", ".fq.bz2", ".fastq.bz2")
    if not os.path.isdir(folder):
        return
    jobs = []
    for f in os.listdir(folder):
        path = os.path.join(folder, f)
        if os.path.isdir(path) or not f.endswith(fqext) or f.startswith("Undetermined"):
            continue
        if not matchFlag(f, options.read1_flag):
            continue
        print(f)
        opt = copy.copy(options)
        opt.read1_file = path
        for attr, flag in (("read2_file", options.read2_flag), ("index1_file", options.index1_flag), ("index2_file", options.index2_flag)):
            cand = path.replace(options.read1_flag, flag)
            if os.path.exists(cand):
                setattr(opt, attr, cand)
        if options.barcode_flag in f and parseBool(options.barcode):
            opt.barcode = True
            opt.trim_front = 0
            opt.trim_front2 = 0
        else:
            opt.barcode = False
        jobs.append(opt)
    if not jobs:
        print("no read files to run with, do you call the program correctly?")
        print("see -h for help")
        return
    if (getattr(options, "gpus", 1) or 1) > 1:
        # every job is sharded over the box's GPUs by itself: one after the other, not all at once
        for o in jobs:
            processOptions(o)
        return
    # the reference starts one process per R1 file (after.py:168-171); here job i runs on GPU i mod (GPUs of the box)
    ngpu = device_count()
    for i, o in enumerate(jobs):
        o.device = (i % ngpu) if ngpu > 0 else -1
    procs = [Process(target=processOptions, args=(o,)) for o in jobs]
    for p in procs:
        p.start()
    for p in procs:
        p.join()
    if any(p.exitcode != 0 for p in procs):
        raise RuntimeError("a directory-mode job failed")


def main(argv=None):
    time1 = time.time()
    (options, _args) = parseCommand(argv)
    normalize_options(options)
    if options.input_dir is None and options.read1_file is None:
        print('specify current dir as input dir')
        options.input_dir = "."
    if options.input_dir is not None:
        if options.debubble:
            print('debubble is outside the scope of the B200 engine, skipping it')
            options.debubble = False
        processDir(options.input_dir, options)
    else:
        if options.barcode_flag in options.read1_file and parseBool(options.barcode):
            options.barcode = True
            options.trim_front = 0
            options.trim_front2 = 0
        else:
            options.barcode = False
        processOptions(options)
    time2 = time.time()
    print('Time used: ' + str(time2 - time1))


if __name__ == "__main__":
    main()
