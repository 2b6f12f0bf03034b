"""Host mirror of QualityControl (qualitycontrol.py:31-157,324-408): everything that is NOT the
per-read loop.  The integer per-cycle counters come from the device (aqc_get_qc /
aqc_get_kmer_*); the O(readLen) float derivations, autoTrim and the k-mer ranking are done
here in python floats so the JSON text is bit-identical to the reference's.
"""
import numpy as np

from . import _abi

MAX_LEN = _abi.MAX_LEN
ALL_BASES = _abi.ALL_BASES
KMER_TOP = 10


def _kmer_dense_str(idx, k):
    s = []
    for j in range(k):
        s.append("ACGT"[(idx >> (2 * (k - 1 - j))) & 3])
    return "".join(s)


def _kmer_side_str(key, k):
    return bytes((int(key) >> (8 * (k - 1 - j))) & 0xFF for j in range(k)).decode("latin-1")


class QualityControl:
    """Counters of one QC slot + the derived statistics the JSON/report consume."""

    def __init__(self, qc_sample=1000000, qc_kmer=8):
        self.sampleLimit = qc_sample
        self.kmerLen = qc_kmer
        self.readLen = 0
        self.baseCounts = {b: [0] * MAX_LEN for b in ALL_BASES}
        self.baseTotalQual = {b: [0] * MAX_LEN for b in ALL_BASES}
        self.percents = {b: [0.0] * MAX_LEN for b in ALL_BASES}
        self.baseMeanQual = {b: [0.0] * MAX_LEN for b in ALL_BASES}
        self.totalQual = [0] * MAX_LEN
        self.totalNum = [0] * MAX_LEN
        self.meanQual = [0.0] * MAX_LEN
        self.gcPercents = [0.0] * MAX_LEN
        self.gcHistogram = [0] * MAX_LEN
        self.meanDiscontinuity = [0.0] * MAX_LEN
        self.totalDiscontinuity = [0.0] * MAX_LEN
        self.topKmerCount = []
        self.totalKmer = 0
        self._kmers = None
        self._sorted = None
        self.gcHistogramFull = [0] * (MAX_LEN + 1)

    # ---- load the integer counters fetched from the engine ---------------------------------
    def load(self, counters, kmers):
        """counters: one _abi.QC_DTYPE record; kmers: (dense_counts, dense_first, side_keys, side_counts, side_first)."""
        self.totalNum = counters["totalNum"].tolist()
        self.totalQual = counters["totalQual"].tolist()
        self.totalDiscontinuity = [float(x) for x in counters["totalDiscontinuity"].tolist()]
        self.gcHistogramFull = counters["gcHistogram"].tolist()
        self.gcHistogram = self.gcHistogramFull[:MAX_LEN]
        for i, b in enumerate(ALL_BASES):
            self.baseCounts[b] = counters["baseCounts"][i].tolist()
            self.baseTotalQual[b] = counters["baseTotalQual"][i].tolist()
        self.totalKmer = int(counters["totalKmer"])
        self._kmers = kmers
        return self

    # ---- derived per-cycle statistics (qualitycontrol.py:124-153) ---------------------------
    # numpy float64 division of exactly-converted integers is the same IEEE operation as the reference's
    # float(a) / float(b), so the JSON text is identical; .tolist() yields python floats.
    def _count_matrix(self):
        return np.array([self.baseCounts[b] for b in ALL_BASES], dtype=np.int64)        # [4][MAX_LEN], A,T,C,G

    def calcReadLen(self):
        """first cycle without any A/T/C/G observation; stays 0 when all MAX_LEN cycles are populated (:124-132)"""
        empty = np.flatnonzero(self._count_matrix().sum(axis=0) == 0)
        if len(empty):
            self.readLen = int(empty[0])

    def calcPercents(self):
        n = self.readLen
        if n == 0:
            return
        cnt = self._count_matrix()[:, :n]
        total = cnt.sum(axis=0).astype(np.float64)
        for i, base in enumerate(ALL_BASES):
            self.percents[base][:n] = (cnt[i].astype(np.float64) / total).tolist()
        g, c = ALL_BASES.index('G'), ALL_BASES.index('C')
        self.gcPercents[:n] = ((cnt[g] + cnt[c]).astype(np.float64) / total).tolist()

    def calcQualities(self):
        n = self.readLen
        if n == 0:
            return
        num = np.array(self.totalNum[:n], dtype=np.int64).astype(np.float64)
        self.meanQual[:n] = (np.array(self.totalQual[:n], dtype=np.int64).astype(np.float64) / num).tolist()
        cnt = self._count_matrix()[:, :n]
        for i, base in enumerate(ALL_BASES):
            q = np.array(self.baseTotalQual[base][:n], dtype=np.int64).astype(np.float64)
            seen = cnt[i] > 0                                       # cycles where the base never occurs keep 0.0 (:148)
            out = np.zeros(n, dtype=np.float64)
            out[seen] = q[seen] / cnt[i][seen].astype(np.float64)
            self.baseMeanQual[base][:n] = out.tolist()

    def calcDiscontinuity(self):
        n = self.readLen
        if n == 0:
            return
        num = np.array(self.totalNum[:n], dtype=np.int64).astype(np.float64)
        self.meanDiscontinuity[:n] = (np.array(self.totalDiscontinuity[:n], dtype=np.float64) / num).tolist()

    def sortKmer(self):
        """sorted(kmerCount.items(), key=count, reverse=True) over an insertion-ordered dict
        (qualitycontrol.py:155-156, quirk Q12): count descending, first-seen ascending."""
        if self._kmers is None:
            self.topKmerCount = []
            return
        dcnt, dfirst, skeys, scnt, sfirst = self._kmers
        k = self.kmerLen
        present = np.flatnonzero(dfirst != np.uint64(_abi.KMER_NEVER))
        cnt = np.concatenate([dcnt[present], scnt]).astype(np.uint64)
        first = np.concatenate([dfirst[present], sfirst]).astype(np.uint64)
        order = np.lexsort((first, np.iinfo(np.uint64).max - cnt))
        nd = len(present)
        self._sorted = (order, present, skeys, cnt, nd)
        top = []
        for j in order[:max(KMER_TOP, 10)]:
            j = int(j)
            name = _kmer_dense_str(int(present[j]), k) if j < nd else _kmer_side_str(skeys[j - nd], k)
            top.append((name, int(cnt[j])))
        self.topKmerCount = top

    def strand_bias_points(self, max_points=1000):
        """(forward, reverse) k-mer counts sampled along the sorted k-mer list, as strandBiasPlotly does
        (qualitycontrol.py:238-257): skip the top min(50, n/2) k-mers, at most 1000 evenly spaced points."""
        if self._sorted is None:
            return [], []
        order, present, skeys, cnt, nd = self._sorted
        n = len(order)
        shift = min(50, n // 2)
        top = min(n - shift, max_points)
        if top <= 0:
            return [], []
        step = max(1, (n - shift) // top)
        k = self.kmerLen
        dense_pos = {int(p): j for j, p in enumerate(present)}          # natural dense index -> entry
        side_pos = {int(key): nd + j for j, key in enumerate(skeys)}
        comp = {65: 84, 84: 65, 67: 71, 71: 67, 97: 116, 116: 97, 99: 103, 103: 99, 78: 78, 10: 10}
        fwd, rev = [], []
        for i in range(top):
            idx = i * step + shift
            if idx >= n:
                break
            j = int(order[idx])
            fwd.append(int(cnt[j]))
            if j < nd:
                d = int(present[j]); r = 0
                for _ in range(k):
                    r = (r << 2) | (3 - (d & 3)); d >>= 2
                jr = dense_pos.get(r)
            else:
                key = int(skeys[j - nd]); r = 0
                for t in range(k):
                    r |= comp.get((key >> (8 * (k - 1 - t))) & 0xFF, 78) << (8 * t)
                jr = side_pos.get(r)
            rev.append(int(cnt[jr]) if jr is not None else 0)
        return fwd, rev

    def qc(self):
        self.calcReadLen()
        self.calcPercents()
        self.calcQualities()
        self.calcDiscontinuity()
        self.sortKmer()

    def squeeze(self):
        """qualitycontrol.py:59-71"""
        n = self.readLen
        self.totalQual = self.totalQual[0:n]
        self.totalNum = self.totalNum[0:n]
        self.meanQual = self.meanQual[0:n]
        self.gcPercents = self.gcPercents[0:n]
        self.gcHistogram = self.gcHistogram[0:n]
        self.meanDiscontinuity = self.meanDiscontinuity[0:n]
        self.totalDiscontinuity = self.totalDiscontinuity[0:n]
        for base in ALL_BASES:
            self.baseCounts[base] = self.baseCounts[base][0:n]
            self.percents[base] = self.percents[base][0:n]
            self.baseMeanQual[base] = self.baseMeanQual[base][0:n]
            self.baseTotalQual[base] = self.baseTotalQual[base][0:n]

    # ---- automatic trimming (qualitycontrol.py:359-408) ------------------------------------------
    # Walk outwards from the middle cycle; the first abnormal cycle on each side bounds the good segment.
    _BASE_RANGE = (0.15, 0.4)      # BASE_BOTTOM, BASE_TOP
    _GC_RANGE = (0.3, 0.7)         # GC_BOTTOM, GC_TOP
    _QUAL_BOTTOM = 20.0

    def isAbnormalCycle(self, this_cycle, comp_cycle, percent_change_threshold):
        gc = self.gcPercents[this_cycle]
        if gc > self._GC_RANGE[1] or gc < self._GC_RANGE[0]:
            return True
        for base in ALL_BASES:
            here = self.percents[base][this_cycle]
            if here > self._BASE_RANGE[1] or here < self._BASE_RANGE[0]:
                return True
            if abs(here - self.percents[base][comp_cycle]) > percent_change_threshold:
                return True
            if self.baseMeanQual[base][this_cycle] < self._QUAL_BOTTOM:
                return True
        return False

    def autoTrim(self):
        n = self.readLen
        center = int(n / 2)
        trimFront = trimTail = 0
        for front in range(center - 1, -1, -1):                     # towards the head, compared with the next cycle
            if self.isAbnormalCycle(front, front + 1, 0.10):
                trimFront = front + 1
                break
        for tail in range(center + 1, n):                           # towards the tail, compared with the previous cycle
            if self.isAbnormalCycle(tail, tail - 1, 0.05):
                trimTail = n - tail
                break
        return (min(int(n * 0.1), trimFront), min(int(n * 0.05), trimTail))
