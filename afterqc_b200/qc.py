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

    # ---- qualitycontrol.py:124-156 ---------------------------------------------------------
    def calcReadLen(self):
        for pos in range(MAX_LEN):
            hasData = False
            for base in ALL_BASES:
                if self.baseCounts[base][pos] > 0:
                    hasData = True
            if not hasData:
                self.readLen = pos
                break

    def calcPercents(self):
        for pos in range(self.readLen):
            total = 0
            for base in ALL_BASES:
                total += self.baseCounts[base][pos]
            for base in ALL_BASES:
                self.percents[base][pos] = float(self.baseCounts[base][pos]) / float(total)
            self.gcPercents[pos] = float(self.baseCounts['G'][pos] + self.baseCounts['C'][pos]) / float(total)

    def calcQualities(self):
        for pos in range(self.readLen):
            self.meanQual[pos] = float(self.totalQual[pos]) / float(self.totalNum[pos])
            for base in ALL_BASES:
                if self.baseCounts[base][pos] > 0:
                    self.baseMeanQual[base][pos] = float(self.baseTotalQual[base][pos]) / float(self.baseCounts[base][pos])

    def calcDiscontinuity(self):
        for pos in range(self.readLen):
            self.meanDiscontinuity[pos] = float(self.totalDiscontinuity[pos]) / float(self.totalNum[pos])

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

    # ---- qualitycontrol.py:359-408 ---------------------------------------------------------
    def autoTrim(self):
        center = int(self.readLen / 2)
        front = center
        tail = center
        bad_in_front = False
        bad_in_tail = False
        for front in range(0, center)[::-1]:
            if self.isAbnormalCycle(front, front + 1, 0.10):
                bad_in_front = True
                break
        for tail in range(center + 1, self.readLen):
            if self.isAbnormalCycle(tail, tail - 1, 0.05):
                bad_in_tail = True
                break
        trimFront = 0
        trimTail = 0
        if bad_in_front:
            trimFront = front + 1
        if bad_in_tail:
            trimTail = self.readLen - tail
        trimFront = min(int(self.readLen * 0.1), trimFront)
        trimTail = min(int(self.readLen * 0.05), trimTail)
        return (trimFront, trimTail)

    def isAbnormalCycle(self, this_cycle, comp_cycle, percent_change_threshold):
        BASE_TOP = 0.4
        BASE_BOTTOM = 0.15
        GC_TOP = 0.7
        GC_BOTTOM = 0.3
        QUAL_BOTTOM = 20.0
        if self.gcPercents[this_cycle] > GC_TOP or self.gcPercents[this_cycle] < GC_BOTTOM:
            return True
        for base in ALL_BASES:
            if self.percents[base][this_cycle] > BASE_TOP or self.percents[base][this_cycle] < BASE_BOTTOM:
                return True
            if abs(self.percents[base][this_cycle] - self.percents[base][comp_cycle]) > percent_change_threshold:
                return True
            if self.baseMeanQual[base][this_cycle] < QUAL_BOTTOM:
                return True
        return False
