"""ctypes mirror of include/afterqc_b200.h (structs, enums, counter indices).

Kept in one place so the product binding (afterqc_b200/_native.py) and the test-only oracle
binding (oracle/oracle.py) marshal exactly the same bytes.
"""
import ctypes as C

import numpy as np

ABI_VERSION = 2
MAX_LEN = 1000          # qualitycontrol.py:23
MAX_KMER = 8
NUM_QC = 4

OK, ERR_INVALID, ERR_CUDA, ERR_TOO_LONG, ERR_KMER_TABLE_FULL, ERR_NOMEM, ERR_TOO_SHORT_STAT = range(7)
ERR_NAMES = {
    ERR_INVALID: "invalid argument", ERR_CUDA: "CUDA failure", ERR_TOO_LONG: "read longer than MAX_LEN",
    ERR_KMER_TABLE_FULL: "non-ACGT k-mer side table full", ERR_NOMEM: "out of memory",
    ERR_TOO_SHORT_STAT: "read of 1..4 bases reached statRead",
}

MEM_HOST, MEM_DEVICE = 0, 1
KERNEL_DEFAULT, KERNEL_WARP, KERNEL_LANE = 0, 1, 2     # aqc_params.filter_kernel
STAT_DEFAULT, STAT_WARP = 0, 1                          # aqc_params.stat_kernel
BATCH_QUAL2_IN_PLACE = 1 << 16                         # aqc_batch.flags

# pair classes, reference priority order (preprocesser.py:436-614)
GOOD, BADTRIM1, BADTRIM2, BADLEN, BADPOL, BADLQC, BADNCT, BADDIFF, BADMISMATCH = range(9)
CLASS_FLAGS = [None, "BADTRIM1", "BADTRIM2", "BADLEN", "BADPOL", "BADLQC", "BADNCT", "BADDIFF", "BADMISMATCH"]

QC_R1_PRE, QC_R2_PRE, QC_R1_POST, QC_R2_POST = range(4)

# scalar counter indices (AQC_C_*)
_C_NAMES = [
    "TOTAL_READS", "TOTAL_BASES_R1", "TOTAL_BASES_R2", "GOOD_READS", "GOOD_BASES_R1", "GOOD_BASES_R2",
    "BADTRIM1", "BADTRIM2", "BADLEN", "BADPOL", "BADLQC", "BADNCT", "BADDIFF", "BADMISMATCH",
    "READ_CORRECTED", "BASE_CORRECTED", "BASE_SKIPPED_CORRECTION", "BASE_ZERO_QUAL_MASKED",
    "OVERLAPPED", "OVERLAP_LEN_SUM", "OVERLAP_BASE_SUM", "OVERLAP_BASE_ERR",
    "TRIMMED_ADAPTER_BASE", "TRIMMED_ADAPTER_READ",
]
CIDX = {n: i for i, n in enumerate(_C_NAMES)}
C_ERR_MATRIX = 32
C_SCALARS = 64
C_OVERLAP_HIST = 64
C_DISTANCE_HIST = 64 + MAX_LEN + 1
C_TOTAL = 64 + 2 * (MAX_LEN + 1)
ALL_BASES = ("A", "T", "C", "G")   # qualitycontrol.py:24

KMER_NEVER = 0xFFFFFFFFFFFFFFFF


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "paired", "trim_front", "trim_tail", "trim_front2", "trim_tail2", "seq_len_req",
        "poly_size_limit", "allow_mismatch_in_poly", "qualified_quality_phred",
        "unqualified_base_limit", "n_base_limit", "no_overlap", "no_correction", "mask_mismatch",
        "qc_sample", "qc_kmer", "kmer_side_log2", "filter_kernel", "stat_kernel")] + [("reserved", C.c_int32 * 5)]

    @classmethod
    def defaults(cls, **kw):
        """after.py:14-93 defaults with trimming resolved to 0 (callers resolve autoTrim first)."""
        p = cls(paired=1, trim_front=0, trim_tail=0, trim_front2=0, trim_tail2=0, seq_len_req=35,
                poly_size_limit=35, allow_mismatch_in_poly=2, qualified_quality_phred=15,
                unqualified_base_limit=60, n_base_limit=5, no_overlap=0, no_correction=0,
                mask_mismatch=0, qc_sample=200000, qc_kmer=8, kmer_side_log2=0, filter_kernel=0, stat_kernel=0)
        for k, v in kw.items():
            if not hasattr(p, k):
                raise KeyError(k)
            setattr(p, k, int(v))
        return p


class Batch(C.Structure):
    _fields_ = [
        ("first_index", C.c_uint64), ("n", C.c_uint32), ("flags", C.c_uint32),
        ("seq1", C.c_void_p), ("qual1", C.c_void_p), ("off1", C.c_void_p),
        ("seq2", C.c_void_p), ("qual2", C.c_void_p), ("off2", C.c_void_p),
    ]


class Records(C.Structure):
    """aqc_records: one parsed batch of the streaming reader (views into the reader's slot buffers)."""
    _fields_ = [
        ("n", C.c_uint64), ("first_index", C.c_uint64), ("slot", C.c_uint32), ("max_len", C.c_uint32),
        ("bytes", C.c_void_p * 4), ("off", C.c_void_p * 4), ("seq_off32", C.c_void_p),
    ]


class Columns(C.Structure):
    """aqc_columns / aqc_columns_out (same layout)"""
    _fields_ = [("names", C.c_void_p), ("name_off", C.c_void_p), ("seqs", C.c_void_p), ("seq_off", C.c_void_p), ("quals", C.c_void_p)]


class Parsed(C.Structure):
    """aqc_parsed: what aqc_fastq_parse_device leaves in HBM (device pointers owned by the context)"""
    _fields_ = [
        ("n_records", C.c_uint64), ("consumed", C.c_uint64), ("bad_record", C.c_uint64), ("seq_bytes", C.c_uint64),
        ("hit_eof", C.c_int32), ("reserved", C.c_int32),
        ("seq", C.c_void_p), ("qual", C.c_void_p), ("off", C.c_void_p), ("line_start", C.c_void_p), ("line_len", C.c_void_p),
        ("text", C.c_void_p),
    ]


HOST_BADBCD1, HOST_BADBCD2 = 16, 17       # host-only pseudo classes of aqc_fastq_emit (barcode pre-pass)

RESULT_DTYPE = np.dtype([
    ("cls", "u1"), ("n_edits", "u1"), ("start1", "<u2"), ("len1", "<u2"), ("start2", "<u2"), ("len2", "<u2"),
    ("ov_offset", "<i2"), ("ov_len", "<u2"), ("ov_diff", "<u2"), ("edits", "<u4", (4,)),
])
assert RESULT_DTYPE.itemsize == 32

OPS_DTYPE = np.dtype([
    ("poly1", "u1"), ("poly2", "u1"), ("lowq1", "<u2"), ("lowq2", "<u2"), ("n1", "<u2"), ("n2", "<u2"),
    ("len1", "<u2"), ("len2", "<u2"), ("ov_offset", "<i2"), ("ov_len", "<u2"), ("ov_diff", "<u2"), ("pad", "u1", (12,)),
])
assert OPS_DTYPE.itemsize == 32

QC_DTYPE = np.dtype([
    ("totalNum", "<i8", (MAX_LEN,)), ("totalQual", "<i8", (MAX_LEN,)),
    ("baseCounts", "<i8", (4, MAX_LEN)), ("baseTotalQual", "<i8", (4, MAX_LEN)),
    ("totalDiscontinuity", "<i8", (MAX_LEN,)), ("gcHistogram", "<i8", (MAX_LEN + 1,)),
    ("totalKmer", "<i8"), ("reads", "<i8"),
])


def edit_fields(e):
    e = int(e)
    return {"pos": e & 0x3FF, "kind": (e >> 10) & 3, "base": (e >> 16) & 0xFF, "qual": (e >> 24) & 0xFF,
            "pos2": (e >> 16) & 0x3FF}


def ptr(a):
    """void* of a numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)
