"""Per-file pipeline driver: the host side of seqFilter.run() (preprocesser.py:234-783).

Everything per-read runs in the engine (afterqc_b200.engine.Engine -> libafterqc_b200.so,
hand-written sm_100a kernels).  This module keeps only what the reference does once per file:
prefilter sampling window (qualitycontrol.py:331-357), autoTrim wiring (:260-280), output
directory layout and naming (:285-371, getMainName :14-17), bad-flag renaming (:206-232),
counter -> JSON assembly (:660-778).

`backend_factory(params)` must return an object with the Engine interface; the default is the
CUDA engine and there is NO CPU fallback (the import fails loudly without the built library).
Tests inject the CPU oracle through the same hook to pin the host logic.
"""
import collections
import json
import os

import numpy as np

from . import _abi, fastq_io
from .qc import QualityControl

READ_TO_SKIP = 1000          # qualitycontrol.py:333
HEAD_ORDER_BASE = 1 << 40    # k-mer order index of the re-stat'd head reads (after the window)


def getMainName(filename):
    """preprocesser.py:14-17"""
    baseName = os.path.basename(filename)
    return baseName.replace(".fastq", "").replace(".fq", "").replace(".gz", "")


def makeDict(opt):
    """the `command` echo of the JSON (preprocesser.py:86-123)"""
    keys = ['index2_flag', 'draw', 'barcode', 'index1_flag', 'seq_len_req', 'index1_file',
            'overlap_output_folder', 'trim_tail', 'trim_pair_same', 'poly_size_limit',
            'good_output_folder', 'debubble_dir', 'index2_file', 'qualified_quality_phred',
            'barcode_flag', 'trim_front', 'barcode_verify', 'read2_file', 'n_base_limit',
            'barcode_length', 'trim_tail2', 'unqualified_base_limit', 'allow_mismatch_in_poly',
            'input_dir', 'read1_file', 'read2_flag', 'store_overlap', 'debubble', 'read1_flag',
            'trim_front2', 'bad_output_folder', 'qc_only', 'qc_sample', 'qc_kmer']
    return {k: getattr(opt, k) for k in keys}


def params_from_options(opt, paired):
    return _abi.Params.defaults(
        paired=1 if paired else 0,
        trim_front=max(opt.trim_front, 0), trim_tail=max(opt.trim_tail, 0),
        trim_front2=max(opt.trim_front2, 0), trim_tail2=max(opt.trim_tail2, 0),
        seq_len_req=opt.seq_len_req, poly_size_limit=opt.poly_size_limit,
        allow_mismatch_in_poly=opt.allow_mismatch_in_poly,
        qualified_quality_phred=opt.qualified_quality_phred,
        unqualified_base_limit=opt.unqualified_base_limit, n_base_limit=opt.n_base_limit,
        no_overlap=1 if opt.no_overlap else 0, no_correction=1 if opt.no_correction else 0,
        mask_mismatch=1 if opt.mask_mismatch else 0, qc_sample=opt.qc_sample, qc_kmer=opt.qc_kmer)


def default_backend(params):
    from .engine import Engine   # fails loudly when the CUDA library is missing
    return Engine(params)


def prefilter_stat(backend, rec, slot, sample_limit, batch_records, shard=(0, 1)):
    """QualityControl.statFile (qualitycontrol.py:331-357) for one file already parsed into `rec`.

    Window = records [999, 999+limit) (all if limit <= 0); if fewer than 1000 reads were counted in
    the window loop, the first 999 are stat'd afterwards."""
    n = rec.n
    s_lo, s_hi = (n * shard[0]) // shard[1], (n * (shard[0] + 1)) // shard[1]   # this rank's records
    lo = READ_TO_SKIP - 1
    if sample_limit > 0:
        hi = min(n, lo + sample_limit)
        stat_reads_num = min(max(n - lo, 0), sample_limit + 1)
    else:
        hi = n
        stat_reads_num = max(n - lo, 0)
    for a in range(max(lo, s_lo), min(hi, s_hi), batch_records):
        b = min(hi, s_hi, a + batch_records)
        batch = fastq_io.to_batch(rec, None, a, b)
        backend.stat_reads(batch, slot, -1, stat_lo=lo, stat_hi=hi, order_base=0)
    if stat_reads_num < READ_TO_SKIP:
        head = min(n, READ_TO_SKIP - 1)
        a, b = max(0, s_lo), min(head, s_hi)
        if b > a:
            batch = fastq_io.to_batch(rec, None, a, b)
            backend.stat_reads(batch, slot, -1, stat_lo=0, stat_hi=head, order_base=HEAD_ORDER_BASE)


def _emit(out, name, seq, plus, qual):
    out.append(name); out.append(b"\n"); out.append(seq); out.append(b"\n")
    out.append(plus); out.append(b"\n"); out.append(qual); out.append(b"\n")


def apply_result(r, s1, q1, s2, q2):
    """Return the final (seq1, qual1, seq2, qual2) bytes of one pair given its aqc_result."""
    n_edits = int(r["n_edits"])
    if n_edits:
        s1 = bytearray(s1); q1 = bytearray(q1)
        if s2 is not None:
            s2 = bytearray(s2); q2 = bytearray(q2)
        for e in r["edits"][:n_edits]:
            f = _abi.edit_fields(e)
            if f["kind"] == 0:
                s1[f["pos"]] = f["base"]; q1[f["pos"]] = f["qual"]
            elif f["kind"] == 1:
                s2[f["pos"]] = f["base"]; q2[f["pos"]] = f["qual"]
            elif f["kind"] == 2:
                q1[f["pos"]] = 0x21; q2[f["pos2"]] = 0x21
    a1, l1 = int(r["start1"]), int(r["len1"])
    o1, p1 = bytes(s1[a1:a1 + l1]), bytes(q1[a1:a1 + l1])
    if s2 is None:
        return o1, p1, None, None
    a2, l2 = int(r["start2"]), int(r["len2"])
    return o1, p1, bytes(s2[a2:a2 + l2]), bytes(q2[a2:a2 + l2])


class _OutputLane:
    """One output file of the streaming loop: formatting (native emit into a reused buffer) and writing run on the
    lane's own thread, strictly in submission order; the native calls and zlib release the GIL."""

    def __init__(self, writer):
        import concurrent.futures
        self.writer = writer
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=1)
        self._scratch = None

    def submit(self, rec, mate, which, base, res):
        return self._pool.submit(self._job, rec, mate, which, base, res)

    def _job(self, rec, mate, which, base, res):
        data, self._scratch = fastq_io.emit_into(rec, mate, which, base, res, self._scratch)
        if len(data):
            self.writer.write(data)

    def shutdown(self):
        self._pool.shutdown(wait=True)


class seqFilter:
    """Drop-in for preprocesser.seqFilter: seqFilter(options).run()."""

    def __init__(self, opt, backend_factory=None, batch_records=1 << 18, shard=(0, 1)):
        self.options = opt
        self.shard = shard            # (rank, world): this process filters records [n*rank/world, n*(rank+1)/world)
        self.backend_factory = backend_factory or default_backend
        self.batch_records = batch_records
        self.paired = opt.read2_file is not None
        self.stat = None

    def run(self):
        opt = self.options
        if getattr(opt, "debubble", False):
            raise NotImplementedError("--debubble is outside the B200 hot-path scope (SURVEY.md section 2, #12)")
        rank, world = self.shard
        self._bcd = [0, 0]                # BADBCD1, BADBCD2: pairs rejected by the barcode pre-pass (never reach the device)
        self._bcd_bases = [0, 0]          # bases the pre-pass kept from the device (R1, R2): TOTAL_BASES counts them (:416-431)
        if getattr(opt, "barcode", False):
            if world > 1:
                raise NotImplementedError("barcode (UMI) files are processed on one GPU (the pre-pass runs in the streaming loop)")
            if opt.barcode_length < 1:
                raise ValueError("barcode_length=%d is outside the supported domain" % opt.barcode_length)
            opt.trim_front = 0            # no front trim if the sequence is barcoded (preprocesser.py:241-243)
        # The files are streamed in bounded memory (two passes, like the reference).  Shards: one counting pass gives the number
        # of complete records n (the loop ends at the shortest file); rank r then statistics its slice of the prefilter window
        # and filters records [n*r/W, n*(r+1)/W), skipping the rest of the stream without holding it.
        # index reads (-7 / -5) are carried along untouched (preprocesser.py:358-371,422-431)
        idx_files = [(k, getattr(opt, k)) for k in ("index1_file", "index2_file") if getattr(opt, k) is not None]
        self._n_total = None
        self._extra = 0
        if world > 1:
            if opt.qc_only:
                raise NotImplementedError("--qc_only stops at a data-dependent record; run it on one GPU")
            self._n_total, self._extra = self._count_records(idx_files)

        params = params_from_options(opt, self.paired)
        be = self.backend_factory(params)
        self.backend = be

        # ---- prefilter QC (preprocesser.py:247-251) ----
        # (the reader of mate 2 is opened first: its thread parses the window while mate 1 is being worked on)
        ahead = fastq_io.open_stream(opt.read2_file, self.batch_records, slots=2) if self.paired else None
        try:
            self._prefilter_stream(be, opt.read1_file, _abi.QC_R1_PRE)
        except BaseException:
            if ahead is not None:
                ahead.close()
            raise
        if self.paired:
            self._prefilter_stream(be, opt.read2_file, _abi.QC_R2_PRE, stream=ahead)
        self.r1qc_prefilter = QualityControl(opt.qc_sample, opt.qc_kmer).load(be.qc(_abi.QC_R1_PRE), be.kmers(_abi.QC_R1_PRE))
        self.r1qc_prefilter.qc()
        self.r2qc_prefilter = QualityControl(opt.qc_sample, opt.qc_kmer)
        if self.paired:
            self.r2qc_prefilter.load(be.qc(_abi.QC_R2_PRE), be.kmers(_abi.QC_R2_PRE))
            self.r2qc_prefilter.qc()

        readLen = self.r1qc_prefilter.readLen

        # ---- auto trim (preprocesser.py:260-280) ----
        if opt.trim_front == -1 or opt.trim_tail == -1:
            trimFront, trimTail = self.r1qc_prefilter.autoTrim()
            if opt.trim_front == -1:
                opt.trim_front = trimFront
            if opt.trim_tail == -1:
                opt.trim_tail = trimTail
            if self.paired:
                if opt.trim_pair_same:
                    opt.trim_front2 = opt.trim_front
                    opt.trim_tail2 = opt.trim_tail
                else:
                    trimFront2, trimTail2 = self.r2qc_prefilter.autoTrim()
                    if opt.trim_front2 == -1:
                        opt.trim_front2 = trimFront2
                    if opt.trim_tail2 == -1:
                        opt.trim_tail2 = trimTail2
        for k in ("trim_front", "trim_tail") + (("trim_front2", "trim_tail2") if self.paired else ()):
            if getattr(opt, k) < 0:
                raise ValueError("%s=%d is outside the supported domain" % (k, getattr(opt, k)))

        if rank == 0:
            print(opt.read1_file + " options:")
            print(opt)

        # ---- output layout (preprocesser.py:285-371) ----
        good_dir = opt.good_output_folder
        if good_dir is None:
            good_dir = os.path.dirname(opt.read1_file)
        bad_dir = opt.bad_output_folder
        if bad_dir is None:
            bad_dir = os.path.join(os.path.dirname(os.path.dirname(good_dir + "/")), "bad")
        overlap_dir = opt.overlap_output_folder
        if overlap_dir is None:
            overlap_dir = os.path.join(os.path.dirname(os.path.dirname(good_dir + "/")), "overlap")
        qc_dir = opt.report_output_folder
        if qc_dir is None:
            qc_dir = os.path.join(os.path.dirname(os.path.dirname(good_dir + "/")), "QC")
        for d in (qc_dir, good_dir, bad_dir):
            os.makedirs(d, exist_ok=True)
        if opt.store_overlap and self.paired:
            os.makedirs(overlap_dir, exist_ok=True)
        gzip_out = opt.gzip or opt.read1_file.endswith(".gz")
        comp = opt.compression

        writers = {}
        if not opt.qc_only:
            part = "" if world == 1 else ".part%04d" % rank   # per-rank pieces, concatenated by rank 0 below

            def mk(d, f, suffix):
                w = fastq_io.Writer(os.path.join(d, getMainName(f) + suffix + part), gzip_out, comp)
                return w
            writers["good1"] = mk(good_dir, opt.read1_file, ".good.fq")
            writers["bad1"] = mk(bad_dir, opt.read1_file, ".bad.fq")
            if opt.store_overlap:
                writers["ov1"] = mk(overlap_dir, opt.read1_file, ".overlap.fq")
            if self.paired:
                writers["good2"] = mk(good_dir, opt.read2_file, ".good.fq")
                writers["bad2"] = mk(bad_dir, opt.read2_file, ".bad.fq")
                if opt.store_overlap:
                    writers["ov2"] = mk(overlap_dir, opt.read2_file, ".overlap.fq")
            for k, f in idx_files:
                writers["good_" + k] = mk(good_dir, f, ".good.fq")
                writers["bad_" + k] = mk(bad_dir, f, ".bad.fq")
                if opt.store_overlap and self.paired:
                    writers["ov_" + k] = mk(overlap_dir, f, ".overlap.fq")

        # ---- the per-read loop, in batches (preprocesser.py:411-631) ----
        params = params_from_options(opt, self.paired)
        be.set_params(params)
        if world == 1 and self._device_parse_applies(be, idx_files):
            extra = self._filter_stream_device_parse(be, writers)
        elif world == 1:
            extra = self._filter_stream(be, writers, idx_files)
        else:
            n = self._n_total
            if opt.index2_file is not None and not self.paired and n > 0:
                raise TypeError("'NoneType' object is not subscriptable")       # preprocesser.py:426-431 without a read2 file
            self._filter_stream(be, writers, idx_files, (n * rank) // world, (n * (rank + 1)) // world)
            # the R1 record read just before a shorter mate/index file ran out is still counted (preprocesser.py:416-431)
            extra = self._extra
        for w in writers.values():
            w.close()
        if world > 1 and not opt.qc_only:
            import torch.distributed as dist
            dist.barrier()
            if rank == 0:
                from .multigpu import concat_outputs
                for w in writers.values():
                    base = w.filename[:-len(".part0000")]
                    final = base
                    # Writer appended ".gz" after the part suffix when gzip is forced: normalise the final name
                    if w.filename.endswith(".gz") and not base.endswith(".gz"):
                        base = w.filename[:-len(".part0000.gz")]
                        final = base + ".gz"
                    pieces = [w.filename.replace(".part0000", ".part%04d" % r) for r in range(world)]
                    concat_outputs(pieces, final)
                    for pth in pieces:
                        os.unlink(pth)

        cnt = be.counters()
        if getattr(opt, "barcode", False):
            cnt = cnt.copy()
            cnt[_abi.CIDX["TOTAL_READS"]] += self._bcd[0] + self._bcd[1]
            cnt[_abi.CIDX["TOTAL_BASES_R1"]] += self._bcd_bases[0]
            cnt[_abi.CIDX["TOTAL_BASES_R2"]] += self._bcd_bases[1]
        self.counters = cnt
        hist_len = readLen + 1
        if cnt[_abi.C_OVERLAP_HIST + hist_len:_abi.C_OVERLAP_HIST + _abi.MAX_LEN + 1].any() or \
           cnt[_abi.C_DISTANCE_HIST + hist_len:_abi.C_DISTANCE_HIST + _abi.MAX_LEN + 1].any():
            raise IndexError("list index out of range")   # overlap_histgram is readLen+1 long (preprocesser.py:257,517)

        self.r1qc_postfilter = QualityControl(opt.qc_sample, opt.qc_kmer).load(be.qc(_abi.QC_R1_POST), be.kmers(_abi.QC_R1_POST))
        self.r1qc_postfilter.qc()
        self.r2qc_postfilter = QualityControl(opt.qc_sample, opt.qc_kmer)
        if self.paired:
            self.r2qc_postfilter.load(be.qc(_abi.QC_R2_POST), be.kmers(_abi.QC_R2_POST))
            self.r2qc_postfilter.qc()

        # quirk Q1: the reference only adds R2's bases when an index2 file is present (preprocesser.py:426-431,622-623)
        self._count_r2_bases = opt.index2_file is not None
        figure_qcs = [("Read1" if self.paired else "", "before", "r1_pre", self.r1qc_prefilter),
                      ("Read1" if self.paired else "", "after", "r1_post", self.r1qc_postfilter)]
        if self.paired:
            figure_qcs += [("Read2", "before", "r2_pre", self.r2qc_prefilter), ("Read2", "after", "r2_post", self.r2qc_postfilter)]
        read_figs = []
        if rank == 0:
            from . import report
            for mate_label, when, key, q in figure_qcs:             # before squeeze(): the GC plot needs readLen+1 bins
                read_figs += report.read_figures(q, mate_label, when, key)
        stat = self._build_stat(cnt, readLen, extra)
        self.stat = stat
        if rank == 0:
            with open(os.path.join(qc_dir, os.path.basename(opt.read1_file) + ".json"), "w") as f:
                f.write(json.dumps(stat, sort_keys=True, indent=4, separators=(',', ': ')))
            self._write_html(os.path.join(qc_dir, os.path.basename(opt.read1_file) + ".html"), stat, cnt, read_figs)
        be.close()
        return stat

    # ------------------------------------------------------------------------------------------
    def _write_html(self, path, stat, cnt, read_figs):
        """the report of preprocesser.py:678-700,771-783 (figure order as the reference)"""
        from . import report
        opt = self.options
        c = lambda name: int(cnt[_abi.CIDX[name]])
        total = c("TOTAL_READS")
        labels = ['good reads', 'has_polyX', 'low_quality', 'too_short', 'too_many_N']
        counts = [c("GOOD_READS"), c("BADPOL"), c("BADLQC"), c("BADLEN") + c("BADTRIM1") + c("BADTRIM2"), c("BADNCT")]
        if self.paired:
            labels.append('bad_overlap'); counts.append(c("BADMISMATCH") + c("BADDIFF"))
        if getattr(opt, "barcode", False):
            labels.append('bad_barcode'); counts.append(self._bcd[0] + self._bcd[1])
        labels = ["%s: %d(%s%%)" % (l, n, (100.0 * float(n) / total) if total > 0 else 0.0) for l, n in zip(labels, counts)]
        figs = [report.filter_figure(labels, counts, total)]
        if self.paired:
            figs.append(report.error_figure(stat["afterqc_overlap"]["error_matrix"]))
            figs.append(report.overlap_figure(self.overlap_histgram, stat["afterqc_main_summary"]["readlen"], total))
        figs += read_figs
        report.write_html(path, stat, getattr(opt, "version", "0.9.6"), figs)

    def _count_records(self, idx_files):
        """Sharded runs: the number of records the per-read loop will see (it ends at the shortest input,
        preprocesser.py:411-431) and the length of the R1 record the reference reads -- and counts in total_bases -- just
        before a shorter mate / index file runs out.  One streaming pass per file, nothing is kept."""
        opt = self.options

        def count(path, want_len_at=None):
            stream = fastq_io.open_stream(path, self.batch_records, slots=2)
            n, extra = 0, 0
            try:
                while True:
                    k = stream.available(self.batch_records)
                    if k == 0:
                        break
                    rec = stream.take(k)
                    if want_len_at is not None and n <= want_len_at < n + rec.n:
                        extra = int(rec.lengths()[want_len_at - n])
                    n += rec.n
                    rec.done()
            finally:
                stream.close()
            self.__dict__.setdefault("_n_cache", {})[path] = n
            return n, extra
        others = ([opt.read2_file] if self.paired else []) + [f for _k, f in idx_files]
        n_other = min([count(f)[0] for f in others]) if others else None
        n1, extra = count(opt.read1_file, n_other)
        if n_other is None or n1 <= n_other:
            return n1, 0
        return n_other, extra

    def _prefilter_stream(self, be, path, slot, stream=None):
        """QualityControl.statFile (qualitycontrol.py:331-357) on a stream: window = records [999, 999+limit) (all when
        limit <= 0); reading stops one record past the window (that is all statFile's counter needs); if fewer than
        1000 records were counted in the window loop the first 999 are stat'd afterwards.  Sharded runs: every rank streams
        the window and statistics an equal slice of it (counters add up in the all-reduce, k-mer stamps carry global order)."""
        opt = self.options
        rank, world = self.shard
        limit = opt.qc_sample
        lo = READ_TO_SKIP - 1
        hi = lo + limit if limit > 0 else None
        # this rank's share [m_lo, m_hi) of the window [lo, w_end) and of the head [0, lo)
        if world > 1:
            w_end = min(self._n_file(path), hi) if hi is not None else self._n_file(path)
            w_end = max(w_end, lo)
            m_lo, m_hi = lo + ((w_end - lo) * rank) // world, lo + ((w_end - lo) * (rank + 1)) // world
        else:
            m_lo, m_hi = lo, None
        if stream is None:
            stream = fastq_io.open_stream(path, self.batch_records, slots=2)  # the window is short: little read-ahead
        g = 0
        head = []
        try:
            while True:
                k = stream.available(self.batch_records)
                if k == 0:
                    break
                rec = stream.take(k)
                a, b = g, g + rec.n
                if a < lo:
                    head.append(rec.slice(0, min(rec.n, lo - a)))
                wa, wb = max(a, lo, m_lo), (b if hi is None else min(b, hi))
                if m_hi is not None:
                    wb = min(wb, m_hi)
                if wb > wa:
                    batch = fastq_io.to_batch(rec, None, wa - a, wb - a, first_index=wa)
                    be.stat_reads(batch, slot, -1, stat_lo=lo, stat_hi=(hi if hi is not None else 1 << 62), order_base=0)
                rec.done()
                g = b
                if hi is not None and g > hi:
                    break
        finally:
            stream.close()
        stat_reads_num = min(max(g - lo, 0), limit + 1) if limit > 0 else max(g - lo, 0)
        if stat_reads_num < READ_TO_SKIP and head:
            hrec = fastq_io.FastqRecords.concat(head)
            h_lo, h_hi = (hrec.n * rank) // world, (hrec.n * (rank + 1)) // world
            if h_hi > h_lo:
                batch = fastq_io.to_batch(hrec, None, h_lo, h_hi, first_index=h_lo)
                be.stat_reads(batch, slot, -1, stat_lo=0, stat_hi=hrec.n, order_base=HEAD_ORDER_BASE)

    def _n_file(self, path):
        """records of one input file (sharded runs; cached)"""
        cache = self.__dict__.setdefault("_n_cache", {})
        if path not in cache:
            stream = fastq_io.open_stream(path, self.batch_records, slots=2)
            n = 0
            try:
                while True:
                    k = stream.available(self.batch_records)
                    if k == 0:
                        break
                    rec = stream.take(k)
                    n += rec.n
                    rec.done()
            finally:
                stream.close()
            cache[path] = n
        return cache[path]

    def _filter_stream(self, be, writers, idx_files, lo=0, hi=None):
        """The per-read loop over lock-stepped streams of R1 [, R2] [, I1] [, I2]; ends at the shortest file
        (preprocesser.py:411-431).  Returns the bases of the R1 record that the reference reads (and counts) just before
        another file runs out.  --qc_only: stop after the first GOOD pair whose TOTAL_READS >= qc_sample (:630-631).
        Sharded runs pass their record range [lo, hi): records before it are skipped in the streams, the loop stops at hi.

        Three stages overlap: the native readers parse the next batches on their own threads, this thread runs the
        device call, and every output file formats + (deflates +) writes its text on its own lane, in batch order."""
        opt = self.options
        paths = [opt.read1_file] + ([opt.read2_file] if self.paired else []) + [f for _k, f in idx_files]
        streams = [fastq_io.open_stream(p, self.batch_records) for p in paths]
        lanes = {name: _OutputLane(w) for name, w in writers.items()}
        keys = [k for k, _f in idx_files]
        qs = opt.qc_sample
        g = 0
        stopped = False
        inflight = collections.deque()

        def retire(entry):
            recs, futs = entry
            for f in futs:
                f.result()
            for r in recs:
                r.done()

        try:
            while g < lo:                                   # skip to this rank's first record
                k = min(s.available(min(self.batch_records, lo - g)) for s in streams)
                if k == 0:
                    break
                for s_ in streams:
                    s_.take(k).done()
                g += k
            while not stopped:
                want = self.batch_records
                if opt.qc_only:
                    want = 1 if g + 1 >= qs else min(want, qs - 1 - g)    # single pairs once the stop rule can fire
                if hi is not None:
                    want = min(want, hi - g)
                    if want <= 0:
                        break
                k = min(s.available(want) for s in streams)
                if k == 0:
                    break
                if opt.index2_file is not None and not self.paired:
                    # the reference adds len(r2[1]) whenever an index2 record was read (preprocesser.py:426-431, quirk Q1):
                    # without a read2 file that is None[1] on the first complete record
                    raise TypeError("'NoneType' object is not subscriptable")
                recs = [s.take(k) for s in streams]
                rec1 = recs[0]
                rec2 = recs[1] if self.paired else None
                if getattr(opt, "barcode", False):
                    rec1, rec2, res = self._barcode_batch(be, rec1, rec2, g)
                else:
                    batch = fastq_io.to_batch(rec1, rec2, 0, k, first_index=g)
                    res = be.filter_pairs(batch)
                futs = []
                if not opt.qc_only:
                    futs = self._write_async(lanes, rec1, rec2, res)
                    for key, r in zip(keys, recs[(2 if self.paired else 1):]):
                        futs.append(lanes["good_" + key].submit(r, 0, 0, 0, res))
                        futs.append(lanes["bad_" + key].submit(r, 0, 1, 0, res))
                        if "ov_" + key in lanes:
                            futs.append(lanes["ov_" + key].submit(r, 0, 2, 0, res))
                inflight.append((recs, futs))
                while len(inflight) > 2:
                    retire(inflight.popleft())
                g += k
                if opt.qc_only and g >= qs and int(res["cls"][-1]) == _abi.GOOD:
                    stopped = True
            while inflight:
                retire(inflight.popleft())
            extra = 0
            if hi is None and not stopped and len(streams) > 1 and streams[0].available(1) > 0:
                extra = int(streams[0].take(1).lengths()[0])
            return extra
        finally:
            for lane in lanes.values():
                lane.shutdown()
            for s_ in streams:
                s_.close()

    # ---- opt-in: FASTQ text parsed on the device (AQC_DEVICE_PARSE=1) ----
    def _device_parse_applies(self, be, idx_files):
        """FASTQ inputs (plain, .gz, .bz2) on one GPU, no barcode / index files / --qc_only, an engine that can parse."""
        opt = self.options
        if os.environ.get("AQC_DEVICE_PARSE") != "1":
            return False
        return not (idx_files or getattr(opt, "barcode", False) or opt.qc_only or not hasattr(be, "parse_fastq"))

    def _filter_stream_device_parse(self, be, writers, block_bytes=None):
        """_filter_stream with the parser on the device (csrc/aqc_parse.cuh): the host reads blocks of text, the engine turns
        them into the packed columns in HBM (Engine.parse_fastq) and filters them where they lie (ParsedDeviceBatch); only
        the 32-byte records and the line table come back, and the writers format from the text the host already holds.
        Same loop semantics as _filter_stream: lock step over the mates, the shortest file ends the run, returns the bases
        of the R1 record the reference reads just before the other file runs out."""
        from .engine import ParsedDeviceBatch
        opt = self.options
        if block_bytes is None:
            # large enough for the record cap of a step to bind on both mates (then no mate is parsed twice); AQC_DEVICE_PARSE_BLOCK: tests
            block_bytes = int(os.environ.get("AQC_DEVICE_PARSE_BLOCK", 0)) or max(64 << 20, 512 * self.batch_records)
        paths = [opt.read1_file] + ([opt.read2_file] if self.paired else [])
        files = [fastq_io.TextSource(p) for p in paths]          # text only: inflate threads for .gz, no host parser
        lanes = {name: _OutputLane(w) for name, w in writers.items()}
        nm = len(files)
        carry = [np.zeros(0, dtype=np.uint8) for _ in files]        # text not yet consumed
        ended = [False] * nm                                        # nothing more to read from the file
        g = 0
        extra = 0
        inflight = collections.deque()

        def read_more(m):
            new = files[m].read(block_bytes)
            if len(new) < block_bytes:
                ended[m] = True
            if len(new):
                carry[m] = np.concatenate([carry[m], new]) if len(carry[m]) else new

        def parse(m, cap):
            try:
                return be.parse_fastq(carry[m], slot=m, final=ended[m], max_records=cap)
            except ValueError as e:           # the record number in the message counts from this block
                raise ValueError(str(e).replace("FASTQ record ", "FASTQ record %d + " % g))

        try:
            while True:
                for m in range(nm):
                    if not ended[m] and len(carry[m]) < block_bytes:
                        read_more(m)
                ps = [parse(m, self.batch_records) for m in range(nm)]
                k = min(p.n for p in ps)
                if k == 0:
                    # a mate without a complete record: more text may complete one; otherwise its file is over (end of the
                    # file or an empty line) and so is the loop, like the reference's at the shortest file
                    starved = [m for m in range(nm) if ps[m].n == 0 and not ps[m].hit_eof and not ended[m]]
                    if starved:
                        for m in starved:
                            if len(carry[m]) > 16 * block_bytes:
                                raise ValueError("no complete FASTQ record in %d bytes of %s" % (len(carry[m]), paths[m]))
                            read_more(m)
                        continue
                    if nm > 1 and ps[0].n > 0:
                        extra = int(ps[0].fetch_lines()[1][1])      # the R1 record read just before the other file ran out
                    break
                for m in range(nm):
                    if ps[m].n > k:
                        ps[m] = parse(m, k)                         # exactly the records of this step (and their `consumed`)
                db = ParsedDeviceBatch(be, ps[0], ps[1] if nm > 1 else None, first_index=g)
                be.filter_pairs(db)
                res = be.fetch_results(db)
                db.free()
                recs = []
                for m in range(nm):
                    ls, ll = ps[m].fetch_lines()
                    recs.append(fastq_io.TextRecords(carry[m], ls, ll, k))      # keeps this block's text alive for the writers
                inflight.append(self._write_async(lanes, recs[0], recs[1] if nm > 1 else None, res))
                while len(inflight) > 2:
                    for f in inflight.popleft():
                        f.result()
                for m in range(nm):
                    carry[m] = carry[m][ps[m].consumed:]
                g += k
            while inflight:
                for f in inflight.popleft():
                    f.result()
            return extra
        finally:
            for lane in lanes.values():
                lane.shutdown()
            for f in files:
                f.close()

    def _barcode_batch(self, be, rec1, rec2, g):
        """Barcode (UMI) files: the pre-pass of preprocesser.py:435-452 on one batch, then the device loop on the pairs
        that kept a barcode.  Returns the transformed records (input order, barcodes in the names) and one result per
        input pair; BADBCD pairs carry a host-only class and their untouched reads.  The device sees compacted batches
        whose first_index keeps the qc_sample gate (:624, TOTAL_READS counts the BADBCD pairs too) and the k-mer order
        of the original indices: pairs before the gate run with the batch's own first index (compaction only lowers an
        index), pairs at or beyond it with the original index of their first pair."""
        from .batch import PackedBatch
        opt = self.options
        n = rec1.n
        t1, t2, status, removed = fastq_io.barcode_transform(rec1, rec2, opt.barcode_length, opt.barcode_verify)
        bad = status != 0
        self._bcd[0] += int((status == 1).sum())
        self._bcd[1] += int((status == 2).sum())
        len1 = t1.lengths()
        self._bcd_bases[0] += removed[0] + int(len1[bad].sum())
        res = np.zeros(n, dtype=_abi.RESULT_DTYPE)
        res["cls"][status == 1] = _abi.HOST_BADBCD1
        res["cls"][status == 2] = _abi.HOST_BADBCD2
        res["len1"][bad] = len1[bad]
        if t2 is not None:
            len2 = t2.lengths()
            self._bcd_bases[1] += removed[1] + int(len2[bad].sum())
            res["len2"][bad] = len2[bad]
        idx = np.flatnonzero(~bad)
        if len(idx):
            def column(t):
                off = t.seqs.off
                if len(idx) == n:
                    return t.seqs.data, t.quals.data, off.astype(np.uint32)
                lens = (off[1:] - off[:-1])[idx]
                noff = np.zeros(len(idx) + 1, dtype=np.int64)
                np.cumsum(lens, out=noff[1:])
                src = np.arange(int(noff[-1]), dtype=np.int64) + np.repeat(off[:-1][idx] - noff[:-1], lens)
                seq = np.zeros(int(noff[-1]) + 64, dtype=np.uint8); qual = np.zeros(int(noff[-1]) + 64, dtype=np.uint8)
                seq[:len(src)] = t.seqs.data[src]; qual[:len(src)] = t.quals.data[src]
                return seq, qual, noff.astype(np.uint32)
            s1, q1, o1 = column(t1)
            s2 = q2 = o2 = None
            if t2 is not None:
                s2, q2, o2 = column(t2)
            qs = opt.qc_sample
            n_gate = int(np.searchsorted(g + idx + 1, qs, side="left")) if qs > 0 else len(idx)   # pairs with TOTAL_READS < qs
            parts = []
            for lo, hi, first in ((0, n_gate, g), (n_gate, len(idx), g + int(idx[min(n_gate, len(idx) - 1)]))):
                if hi > lo:
                    sub = PackedBatch(s1, q1, o1[lo:hi + 1], s2, q2, (o2[lo:hi + 1] if o2 is not None else None), first_index=first)
                    parts.append(be.filter_pairs(sub))
            res[idx] = np.concatenate(parts) if len(parts) > 1 else parts[0]
        rec1.done()
        if rec2 is not None:
            rec2.done()
        return t1, t2, res

    def _write_async(self, lanes, rec1, rec2, res):
        """_write() on the output lanes; returns the futures"""
        futs = [lanes["good1"].submit(rec1, 1, 0, 0, res), lanes["bad1"].submit(rec1, 1, 1, 0, res)]
        if rec2 is not None:
            futs += [lanes["good2"].submit(rec2, 2, 0, 0, res), lanes["bad2"].submit(rec2, 2, 1, 0, res)]
            if "ov1" in lanes:
                futs += [lanes["ov1"].submit(rec1, 1, 2, 0, res), lanes["ov2"].submit(rec2, 2, 2, 0, res)]
        return futs

    def _write(self, writers, rec1, rec2, base, res):
        """good/bad(/overlap) text of one batch: slices, correction edits and @BADxxx names (preprocesser.py:206-232)."""
        writers["good1"].write(fastq_io.emit(rec1, 1, 0, base, res))
        writers["bad1"].write(fastq_io.emit(rec1, 1, 1, base, res))
        if rec2 is not None:
            writers["good2"].write(fastq_io.emit(rec2, 2, 0, base, res))
            writers["bad2"].write(fastq_io.emit(rec2, 2, 1, base, res))
        if "ov1" in writers:
            if rec2 is not None:
                writers["ov1"].write(fastq_io.emit(rec1, 1, 2, base, res))
                writers["ov2"].write(fastq_io.emit(rec2, 2, 2, base, res))

    def _build_stat(self, cnt, readLen, extra_total_bases):
        """preprocesser.py:660-778"""
        opt = self.options
        c = lambda name: int(cnt[_abi.CIDX[name]])
        result = {
            'total_bases': c("TOTAL_BASES_R1") + extra_total_bases + (c("TOTAL_BASES_R2") if self._count_r2_bases else 0),
            'good_bases': c("GOOD_BASES_R1") + (c("GOOD_BASES_R2") if self._count_r2_bases else 0),
            'total_reads': c("TOTAL_READS"),
            'good_reads': c("GOOD_READS"),
            'bad_reads': c("TOTAL_READS") - c("GOOD_READS"),
            'bad_reads_with_bad_barcode': self._bcd[0] + self._bcd[1],
            'bad_reads_with_reads_in_bubble': 0,
            'bad_reads_with_bad_read_length': c("BADLEN") + c("BADTRIM1") + c("BADTRIM2"),
            'bad_reads_with_polyX': c("BADPOL"),
            'bad_reads_with_low_quality': c("BADLQC"),
            'bad_reads_with_too_many_N': c("BADNCT"),
            'bad_reads_with_bad_overlap': c("BADMISMATCH") + c("BADDIFF"),
            'readlen': readLen,
        }
        qcs = [("read1_prefilter", self.r1qc_prefilter), ("read1_postfilter", self.r1qc_postfilter)]
        if self.paired:
            qcs += [("read2_prefilter", self.r2qc_prefilter), ("read2_postfilter", self.r2qc_postfilter)]
        for _, q in qcs:
            q.squeeze()
        stat = {"afterqc_main_summary": result, "command": makeDict(opt),
                "kmer_content": {}, "base_quality": {}, "mean_quality": {}, "base_content": {}, "gc_content": {}}
        for name, q in qcs:
            stat["kmer_content"][name] = [list(t) for t in q.topKmerCount[0:10]]
            stat["base_quality"][name] = q.baseMeanQual
            stat["mean_quality"][name] = q.meanQual
            stat["base_content"][name] = q.percents
            stat["gc_content"][name] = q.gcPercents
        if self.paired:
            ov = {}
            overlapped = c("OVERLAPPED")
            ov['overlapped_pairs'] = overlapped
            ov['average_overlap_length'] = float(c("OVERLAP_LEN_SUM") // overlapped) if overlapped > 0 else 0.0
            ov['bad_mismatch_reads'] = c("BADMISMATCH")
            ov['bad_diff'] = c("BADDIFF")
            ov['bad_indel_reads'] = 0
            ov['corrected_reads'] = c("READ_CORRECTED")
            ov['corrected_bases'] = c("BASE_CORRECTED")
            ov['skipped_correction_bases'] = c("BASE_SKIPPED_CORRECTION")
            ov['zero_qual_masked'] = c("BASE_ZERO_QUAL_MASKED")
            ov['zero_qual_skipped'] = c("BASE_ZERO_QUAL_MASKED")
            ov['trimmed_adapter_bases'] = c("TRIMMED_ADAPTER_BASE")
            ov['trimmed_adapter_reads'] = c("TRIMMED_ADAPTER_READ")
            base_sum = c("OVERLAP_BASE_SUM")
            ov['error_rate'] = float(c("OVERLAP_BASE_ERR")) / float(base_sum) if base_sum > 0 else 0.0
            em = {}
            for i, cb in enumerate(_abi.ALL_BASES):
                em[cb] = {}
                for j, eb in enumerate(_abi.ALL_BASES):
                    if cb != eb:
                        em[cb][eb] = int(cnt[_abi.C_ERR_MATRIX + i * 4 + j])
            ov['error_matrix'] = em
            dh = cnt[_abi.C_DISTANCE_HIST:_abi.C_DISTANCE_HIST + readLen + 1].tolist()
            ov['edit_distance_histogram'] = dh[0:10]
            stat["afterqc_overlap"] = ov
            self.overlap_histgram = cnt[_abi.C_OVERLAP_HIST:_abi.C_OVERLAP_HIST + readLen + 1].tolist()
        return stat
