"""HTML report (QC/<R1 basename>.html): the reference's report sections rebuilt from the engine's counters.

Same content as qcreporter.py + the *Plotly() methods of qualitycontrol.py (:158-322) -- summary table, filtering pie,
sequencing-error matrix, overlap-length histogram and, per mate and pre/post filtering, quality / base content / GC /
discontinuity curves and the k-mer strand-bias scatter -- but every figure is a plain dict serialised with json.dumps
into Plotly.newPlot(div, data, layout).  Byte parity of the HTML is not a goal (the reference's own text depends on
py2 str(float)); the numbers are the ones of the bit-exact JSON.
"""
import json

import numpy as np

from . import _abi

BASE_COLORS = {'A': 'rgba(255,0,0,0.5)', 'T': 'rgba(128,0,128,0.5)', 'C': 'rgba(0,255,0,0.5)', 'G': 'rgba(0,0,255,0.5)'}
_COMP = {"A": "T", "T": "A", "C": "G", "G": "C", "a": "t", "t": "a", "c": "g", "g": "c", "N": "N", "\n": "\n"}

CSS = """<style type="text/css">
#menu {text-align:left;}
.menu-item{font-size:14px;padding:4px;}
#container {text-align:center;padding-left:30px;}
.figure-title {color:#bbbbbb;font-size:30px;padding:10px;text-align:left;}
.figure-div {margin-top:40px;text-align:center}
.summary-table {padding:5px;border:1px solid #eeeeee;width:800px}
.col1 {text-align:right;padding:5px;padding-right:20px;color:#666666;}
.col2 {text-align:left;padding:5px;padding-left:20px;color:#332299;}
.plotly-div {width:800;height:600;text-align:center;}
li {color:#666666;font-size:15px;border:0px;}
</style>
"""


def anchor(title):
    return title.replace(" ", "-").replace(".", "-").replace("/", "-")


def human(num):
    """qcreporter.formatNumber: 1024-based K/M/G suffixes"""
    num = float(num)
    units = ["", "K", "M", "G", "T", "P"]
    order = 0
    while num > 1024.0:
        order += 1
        num /= 1024.0
    return str(int(num)) if order == 0 else "%0.3f %s" % (num, units[order])


class Figure:
    def __init__(self, title, div, traces, layout):
        self.title, self.div, self.traces, self.layout = title, div, traces, layout

    def script(self):
        return "Plotly.newPlot(%s, %s, %s);" % (json.dumps(self.div), json.dumps(self.traces), json.dumps(self.layout))


def line(x, y, name, color, width=1):
    return {"x": x, "y": y, "name": name, "mode": "lines", "line": {"color": color, "width": width}}


def quality_figure(qc, div, title):
    x = list(range(qc.readLen))
    traces = [line(x, qc.baseMeanQual[b][0:qc.readLen], b, BASE_COLORS[b]) for b in _abi.ALL_BASES]
    traces.append(line(x, qc.meanQual[0:qc.readLen], "mean", "rgba(20,20,20,255)"))
    return Figure(title, div, traces, {"title": title, "xaxis": {"title": "cycles"}, "yaxis": {"title": "quality"}})


def content_figure(qc, div, title):
    x = list(range(qc.readLen))
    traces = [line(x, qc.percents[b][0:qc.readLen], b, BASE_COLORS[b]) for b in _abi.ALL_BASES]
    traces.append(line(x, qc.gcPercents[0:qc.readLen], "GC", "rgba(20,20,20,255)"))
    return Figure(title, div, traces, {"title": title, "xaxis": {"title": "cycles"}, "yaxis": {"title": "percents", "range": [0.0, 0.8]}})


def gc_figure(qc, div, title):
    if qc.readLen == 0:
        return None
    xs = [100.0 * float(t) / qc.readLen for t in range(qc.readLen + 1)]
    hist = list(qc.gcHistogramFull[0:qc.readLen + 1])
    return Figure(title, div, [{"x": xs, "y": hist, "type": "bar"}],
                  {"title": title, "xaxis": {"title": "percents(%)"}, "yaxis": {"title": "counts"}})


def discontinuity_figure(qc, div, title):
    y = qc.meanDiscontinuity[0:qc.readLen]
    top = (max(qc.meanDiscontinuity) if len(qc.meanDiscontinuity) else 0.0) * 1.5
    return Figure(title, div, [line(list(range(qc.readLen)), y, "discontinuity", "rgba(100,150,0,0.5)", 2)],
                  {"title": title, "xaxis": {"title": "cycles"}, "yaxis": {"title": "discontinuity", "range": [0.0, top]}})


def strand_bias_figure(qc, div, title):
    if qc.readLen == 0:
        return None
    fwd, rev = qc.strand_bias_points()
    hi = max([0] + fwd + rev)
    return Figure(title, div, [{"x": fwd, "y": rev, "mode": "markers", "type": "scatter", "marker": {"size": 2, "color": "rgba(0,0,50,128)"}}],
                  {"title": title, "xaxis": {"title": "relative forward strand KMER count", "range": [-10, hi]},
                   "yaxis": {"title": "relative reverse strand KMER count", "range": [-10, hi]}})


def filter_figure(labels, counts, total_reads):
    title = "Filtering statistics of sampled %d reads" % total_reads
    return Figure("Good reads and bad reads after filtering", "filter_stat",
                  [{"values": counts, "labels": labels, "textinfo": "none", "type": "pie"}],
                  {"title": title, "width": 800, "height": 600})


def error_figure(matrix):
    names, values, colors = [], [], []
    transitions = {("A", "G"), ("G", "A"), ("C", "T"), ("T", "C")}
    for c in _abi.ALL_BASES:
        for e in _abi.ALL_BASES:
            if c != e:
                names.append(c + "->" + e)
                values.append(matrix[c][e])
                colors.append("rgba(246, 103, 0,1.0)" if (c, e) in transitions else "rgba(22, 96, 167,1.0)")
    return Figure("Sequence error distribution", "error_matrix", [{"x": names, "y": values, "marker": {"color": colors}, "type": "bar"}],
                  {"title": "sequencing error transform distribution", "xaxis": {"title": "seq error transform"}, "yaxis": {"title": "counts"}})


def overlap_figure(hist, read_len, total_reads):
    pct = int(hist[0] * 100.0 / total_reads) if total_reads > 0 and len(hist) else 0
    return Figure("Overlap length distribution", "overlap_stat", [{"x": list(range(read_len + 1)), "y": list(hist), "type": "bar"}],
                  {"title": "Pair overlap Length Histgram",
                   "xaxis": {"title": "overlap Length (%d%% not overlapped)" % pct, "range": [-2, read_len]}, "yaxis": {"title": "counts"}})


def read_figures(qc, mate_label, when_label, key):
    """the five per-read figures of preprocesser.py:785-819"""
    t = (mate_label + " " if mate_label else "")
    lower = lambda s: s if mate_label else s[0].upper() + s[1:]
    figs = [
        quality_figure(qc, key + "_quality", lower("%squality curve %s filtering" % (t, when_label))),
        content_figure(qc, key + "_content", lower("%sbase content distribution %s filtering" % (t, when_label))),
        gc_figure(qc, key + "_gc", lower("%sGC curve %s filtering" % (t, when_label))),
        discontinuity_figure(qc, key + "_discontinuity", lower("%sper base discontinuity %s filtering" % (t, when_label))),
        strand_bias_figure(qc, key + "_sb", lower("%skmer strand bias %s filtering" % (t, when_label))),
    ]
    return [f for f in figs if f is not None]


def summary_rows(stat, version):
    s = stat["afterqc_main_summary"]
    paired = stat["command"]["read2_file"] is not None
    rows = [("AfterQC Version:", version),
            ("sequencing:", ("2*%d pair end" % s["readlen"]) if paired else ("%d single end" % s["readlen"])),
            ("total reads:", human(s["total_reads"]))]
    tr, tb = float(s["total_reads"]) or 1.0, float(s["total_bases"]) or 1.0
    rows.append(("filtered out reads:", "%0.3f <font color='#aaaaaa'>(%0.3f%%)</font>" % (s["bad_reads"], 100.0 * s["bad_reads"] / tr)))
    rows.append(("total bases:", human(s["total_bases"])))
    fb = s["total_bases"] - s["good_bases"]
    rows.append(("filtered out bases:", "%0.3f <font color='#aaaaaa'>(%0.3f%%)</font>" % (fb, 100.0 * fb / tb)))
    if paired:
        o = stat["afterqc_overlap"]
        rows += [("estimated seq error:", "%0.3f%%" % (o["error_rate"] * 100)),
                 ("adapter trimmed reads:", human(o["trimmed_adapter_reads"])),
                 ("adapter trimmed bases:", human(o["trimmed_adapter_bases"]))]
    rows.append(("auto trimming", "front:%s, tail:%s (use <font color='#aaaaaa'>-f0 -t0</font> to disable)" %
                 (stat["command"]["trim_front"], stat["command"]["trim_tail"])))
    return rows


def write_html(path, stat, version, figures):
    out = ["<HTML>\n<HEAD>\n", '<script src="http://cdn.plot.ly/plotly-latest.min.js"></script>\n', CSS, "</HEAD>\n<BODY>\n<DIV id='container'>\n"]
    out.append("<div id='menu'><ul>\n<li class='menu-item'><a href='#summary'>1, AfterQC summary</a> </li>\n")
    for i, f in enumerate(figures, start=2):
        out.append("<li class='menu-item'><a href='#%s'>%d, %s</a> </li>\n" % (anchor(f.title), i, f.title))
    out.append("</ul></div>\n<div class='figure-div'>\n<div class='figure-title'><a name='summary'>1, AfterQC summary</a></div>\n<table class='summary-table'>\n")
    for k, v in summary_rows(stat, version):
        out.append("<tr><td class='col1'>%s</td><td class='col2'>%s</td></tr>\n" % (k, v))
    out.append("</table>\n</div>\n<div id='figures'>")
    for i, f in enumerate(figures, start=2):
        out.append("<div class='figure-div'>\n<div class='figure-title'><a name='%s'>%d, %s</a></div>\n" % (anchor(f.title), i, f.title))
        out.append("<div id='%s' class='plotly-div'></div>\n<div class='figure-summary'></div>\n</div>\n" % f.div)
    out.append("</div>\n</DIV>\n<script type=\"text/javascript\">\n")
    for f in figures:
        out.append(f.script() + "\n")
    out.append("</script>\n</BODY>\n</HTML>")
    with open(path, "w") as fh:
        fh.write("".join(out))
