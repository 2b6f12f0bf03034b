// aqc_stream.cpp -- streaming FASTQ reader of libafterqc_b200.so: a background thread
// inflates/reads one file and parses it into a ring of reusable packed-column buffers, so that the host loop only
// hands ready batches to the device.  Mirrors fastq.Reader (fastq.py:17-55): .gz by extension, every line rstrip()'d,
// the first empty line ends the file (quirk Q13).  No CUDA here.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/afterqc_b200.h"
#include "aqc_inflate.hpp"
#include "aqc_pinflate.hpp"

namespace {

constexpr size_t kReadBlock = 4u << 20;
constexpr size_t kSlack = 64;
constexpr uint64_t kMaxColumn = (1ull << 32) - 64;       // aqc_batch offsets are uint32

struct Slot {
    uint8_t *bytes[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t cap[4] = {0, 0, 0, 0};
    uint64_t *off[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t *off32 = nullptr;
    uint64_t n = 0, first = 0;
    uint32_t max_len = 0;
};

// Runs a blocking read(dst, cap) source (the gzip decoders) on its own thread, a few blocks ahead of the parser.
class AsyncSource {
  public:
    using ReadFn = std::function<long(uint8_t *, size_t)>;
    AsyncSource(ReadFn fn, size_t block_bytes, int blocks) : fn_(std::move(fn)) {
        blocks_.resize(blocks);
        for (auto &b : blocks_) b.data.resize(block_bytes);
        for (int i = 0; i < blocks; i++) free_.push_back(i);
        th_ = std::thread([this] { run(); });
    }
    ~AsyncSource() {
        { std::unique_lock<std::mutex> lk(mu_); stop_ = true; cv_.notify_all(); }
        if (th_.joinable()) th_.join();
    }
    // next block copied to dst (cap >= block_bytes): bytes, 0 at the end, -1 after a source error
    long get(uint8_t *dst) {
        int id;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return !ready_.empty() || done_; });
            if (ready_.empty()) return failed_ ? -1 : 0;
            id = ready_.front(); ready_.pop_front();
        }
        const long n = blocks_[id].n;
        memcpy(dst, blocks_[id].data.data(), (size_t)n);
        { std::unique_lock<std::mutex> lk(mu_); free_.push_back(id); cv_.notify_all(); }
        return n;
    }

  private:
    struct Block { std::vector<uint8_t> data; long n = 0; };
    void run() {
        for (;;) {
            int id;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !free_.empty(); });
                if (stop_) return;
                id = free_.front(); free_.pop_front();
            }
            long n = fn_(blocks_[id].data.data(), blocks_[id].data.size());
            std::unique_lock<std::mutex> lk(mu_);
            if (n > 0) { blocks_[id].n = n; ready_.push_back(id); }
            else { failed_ = n < 0; done_ = true; }
            cv_.notify_all();
            if (n <= 0) return;
        }
    }
    ReadFn fn_;
    std::vector<Block> blocks_;
    std::deque<int> free_, ready_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::thread th_;
    bool stop_ = false, done_ = false, failed_ = false;
};

// decoder threads per .gz input: AQC_INFLATE_THREADS, else a quarter of the host's hardware threads (3..8, or 1);
// AQC_INFLATE=serial keeps the single-threaded decoder
int inflate_threads() {
    const char *sel = getenv("AQC_INFLATE");
    if (sel && strcmp(sel, "serial") == 0) return 1;
    if (const char *e = getenv("AQC_INFLATE_THREADS")) { int v = atoi(e); return v < 1 ? 1 : (v > 64 ? 64 : v); }
    unsigned hw = std::thread::hardware_concurrency();
    int v = (int)(hw / 4);
    if (v < 3) return 1;                      // the speculative decoder pays off from about 3 threads (16-bit symbols, two passes)
    return v > 8 ? 8 : v;
}

bool ends_with(const std::string &s, const char *suf) {
    size_t k = strlen(suf);
    return s.size() >= k && s.compare(s.size() - k, k, suf) == 0;
}

}  // namespace

struct aqc_reader {
    std::string path, msg;
    uint64_t batch = 0;
    std::vector<Slot> slots;
    std::deque<uint32_t> free_, ready;
    std::mutex mu;
    std::condition_variable cv;
    std::thread th;
    std::atomic<bool> stop{false};
    bool finished = false;
    int err = 0;
    // source (reader thread only)
    gzFile gz = nullptr;                    // zlib path: pipes / non-mappable files, or AQC_INFLATE=zlib
    FILE *fp = nullptr;
    aqc::GzipInflater *inf = nullptr;       // own decoder over the mapped .gz file
    aqc::ParallelGunzip *pinf = nullptr;    // ... multi-threaded for large files
    AsyncSource *async = nullptr;           // gz only: the decoder runs on its own thread, ahead of the parser
    void *map = nullptr;
    size_t map_len = 0;
    int map_fd = -1;
    std::vector<uint8_t> in;
    bool raw = false;                       // aqc_text_open: no parser thread, no slots
    bool plain_map = false;                 // uncompressed regular file: parsed straight from the mapping (no read copy, no memmove)
    const uint8_t *src() const { return plain_map ? (const uint8_t *)map : in.data(); }
    size_t in_pos = 0, in_end = 0;
    bool src_eof = false, file_end = false;
    uint64_t next_index = 0;

    bool grow(Slot &s, int c, size_t need) {
        if (need <= s.cap[c]) return true;
        size_t cap = s.cap[c] ? s.cap[c] : (size_t)1 << 20;
        while (cap < need) cap += cap / 2;
        uint8_t *p = (uint8_t *)realloc(s.bytes[c], cap);
        if (!p) return false;
        s.bytes[c] = p; s.cap[c] = cap;
        return true;
    }

    // one block of decompressed bytes (gz inputs; runs on the AsyncSource thread): -1 + msg on error
    long read_gz(uint8_t *dst, size_t cap) {
        if (pinf) {
            long got = pinf->read(dst, cap);
            if (got < 0) msg = "gzip: " + pinf->error() + " (AQC_INFLATE=zlib selects zlib's decoder)";
            return got;
        }
        if (inf) {
            long got = inf->read(dst, cap);
            if (got < 0) msg = "gzip: " + inf->error() + " (AQC_INFLATE=zlib selects zlib's decoder)";
            return got;
        }
        long got = gzread(gz, dst, (unsigned)cap);
        if (got < 0) { int e; msg = std::string("gzip: ") + gzerror(gz, &e); return -1; }
        if (got == 0) {
            int e = 0; const char *m = gzerror(gz, &e);
            if (e != Z_OK && e != Z_STREAM_END) { msg = std::string("gzip: ") + m; return -1; }
        }
        return got;
    }

    // more input behind the unparsed tail; returns false on a read error
    bool refill() {
        if (plain_map) {                                                // the window just grows over the mapping
            const size_t more = std::min(kReadBlock, map_len - in_end);
            if (more == 0) src_eof = true;
            in_end += more;
            return true;
        }
        if (in_pos > 0) {
            memmove(in.data(), in.data() + in_pos, in_end - in_pos);
            in_end -= in_pos; in_pos = 0;
        }
        if (in.size() < in_end + kReadBlock) in.resize(in_end + kReadBlock);
        long got;
        if (async) {
            got = async->get(in.data() + in_end);
            if (got < 0) return false;                                  // msg was set by the source thread
        } else if (inf) {
            got = inf->read(in.data() + in_end, kReadBlock);
            if (got < 0) { msg = "gzip: " + inf->error() + " (AQC_INFLATE=zlib selects zlib's decoder)"; return false; }
        } else if (gz) {
            got = gzread(gz, in.data() + in_end, (unsigned)kReadBlock);
            if (got < 0) { int e; msg = std::string("gzip: ") + gzerror(gz, &e); return false; }
            if (got == 0) {
                int e = 0; const char *m = gzerror(gz, &e);
                if (e != Z_OK && e != Z_STREAM_END) { msg = std::string("gzip: ") + m; return false; }
            }
        } else {
            got = (long)fread(in.data() + in_end, 1, kReadBlock, fp);
            if (got == 0 && ferror(fp)) { msg = "read error"; return false; }
        }
        if (got == 0) src_eof = true;
        in_end += (size_t)got;
        return true;
    }

    // parse up to `batch` records into s; returns 0 or an AQC_ERR_* code
    int fill(Slot &s) {
        s.n = 0; s.first = next_index; s.max_len = 0;
        for (int c = 0; c < 4; c++) s.off[c][0] = 0;
        while (s.n < batch && !file_end) {
            if (stop.load(std::memory_order_relaxed)) { s.n = 0; return 0; }      // closing: drop the partial batch
            const size_t avail = in_end - in_pos;
            bool progressed = false;
            if (avail > 0 || src_eof) {
                for (int c = 0; c < 4; c++)
                    if (!grow(s, c, (size_t)s.off[c][s.n] + avail + kSlack)) return AQC_ERR_NOMEM;
                uint64_t *offp[4] = {s.off[0] + s.n, s.off[1] + s.n, s.off[2] + s.n, s.off[3] + s.n};
                uint64_t nrec = 0, consumed = 0, bad = 0;
                int hit = 0;
                int rc = aqc_fastq_parse(src() + in_pos, avail, src_eof ? 1 : 0, batch - s.n, s.bytes, offp, &nrec, &consumed, &hit, &bad);
                if (rc) {
                    char t[128];
                    snprintf(t, sizeof t, "FASTQ record %llu: quality line length differs from sequence length",
                             (unsigned long long)(next_index + s.n + bad));
                    msg = t;
                    s.n += nrec;                         // records before the bad one are still delivered
                    return rc;
                }
                s.n += nrec; in_pos += consumed;
                progressed = nrec > 0;
                if (hit) { file_end = true; break; }
                if (s.n >= batch) break;
            }
            if (src_eof) { if (!progressed) file_end = true; continue; }
            if (!refill()) return AQC_ERR_INVALID;
        }
        return 0;
    }

    void finish(Slot &s) {
        if (s.off[1][s.n] > kMaxColumn) { err = AQC_ERR_INVALID; msg = "batch column exceeds uint32 offsets; use a smaller batch"; return; }
        uint32_t m = 0;
        for (uint64_t i = 0; i <= s.n; i++) s.off32[i] = (uint32_t)s.off[1][i];
        for (uint64_t i = 0; i < s.n; i++) { uint32_t l = s.off32[i + 1] - s.off32[i]; if (l > m) m = l; }
        s.max_len = m;
        for (int c = 1; c <= 3; c += 2) {
            if (!grow(s, c, (size_t)s.off[1][s.n] + kSlack)) { err = AQC_ERR_NOMEM; return; }
            memset(s.bytes[c] + s.off[1][s.n], 0, kSlack);
        }
        next_index += s.n;
    }

    void run() {
        for (;;) {
            uint32_t id;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !free_.empty(); });
                if (stop) break;
                id = free_.front(); free_.pop_front();
            }
            Slot &s = slots[id];
            int rc = fill(s);
            if (stop.load()) break;
            int e2 = 0;
            if (s.n) { finish(s); e2 = err; }
            std::unique_lock<std::mutex> lk(mu);
            if (s.n && !e2) ready.push_back(id); else free_.push_back(id);
            if (rc || e2) { err = rc ? rc : e2; finished = true; }
            else if (file_end || s.n == 0) finished = true;
            cv.notify_all();
            if (finished) break;
        }
    }
};

extern "C" {

static int reader_open(const char *path, uint64_t batch_records, uint32_t slots, bool raw, aqc_reader **out);
int aqc_reader_open(const char *path, uint64_t batch_records, uint32_t slots, aqc_reader **out) {
    if (!path || !out || batch_records == 0) return AQC_ERR_INVALID;
    return reader_open(path, batch_records, slots, false, out);
}
// The reader's sources without its parser: decompressed (or plain) text for callers that parse elsewhere (the device parser).
int aqc_text_open(const char *path, aqc_reader **out) {
    if (!path || !out) return AQC_ERR_INVALID;
    return reader_open(path, 1, 2, true, out);
}
// up to cap bytes of text (cap >= 4 MiB); 0 at the end of the file, -1 after an error (aqc_reader_error)
int64_t aqc_text_read(aqc_reader *r, uint8_t *dst, uint64_t cap) {
    if (!r || !dst || !r->raw) return -1;
    uint64_t filled = 0;
    while (!r->src_eof && filled + kReadBlock <= cap) {
        long got;
        if (r->plain_map) {
            got = (long)std::min<size_t>(kReadBlock, r->map_len - r->in_end);
            memcpy(dst + filled, (const uint8_t *)r->map + r->in_end, (size_t)got);
            r->in_end += (size_t)got;
        } else if (r->async) {
            got = r->async->get(dst + filled);
        } else {
            got = (long)fread(dst + filled, 1, kReadBlock, r->fp);
            if (got == 0 && ferror(r->fp)) { r->msg = "read error"; return -1; }
        }
        if (got < 0) return -1;
        if (got == 0) { r->src_eof = true; break; }
        filled += (uint64_t)got;
    }
    return (int64_t)filled;
}
static int reader_open(const char *path, uint64_t batch_records, uint32_t slots, bool raw, aqc_reader **out) {
    if (slots < 2) slots = 2;
    aqc_reader *r = new aqc_reader();
    r->path = path; r->batch = batch_records;
    if (ends_with(r->path, ".gz")) {
        const char *sel = getenv("AQC_INFLATE");
        if (!(sel && strcmp(sel, "zlib") == 0)) {                       // map the file for the own decoder
            int fd = open(path, O_RDONLY);
            struct stat st;
            if (fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
                void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                if (m != MAP_FAILED) {
                    madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                    r->map = m; r->map_len = (size_t)st.st_size; r->map_fd = fd;
                    int nt = inflate_threads();
                    if (nt > 1 && r->map_len >= aqc::ParallelGunzip::kMinSize) r->pinf = new aqc::ParallelGunzip((const uint8_t *)m, r->map_len, nt);
                    else r->inf = new aqc::GzipInflater((const uint8_t *)m, r->map_len);
                }
            }
            if (!r->inf && !r->pinf && fd >= 0) close(fd);
        }
        if (!r->inf && !r->pinf) {
            r->gz = gzopen(path, "rb");
            if (r->gz) gzbuffer(r->gz, 1u << 20);
        }
    } else {
        const char *sel = getenv("AQC_READER_MMAP");
        int fd = (sel && sel[0] == '0') ? -1 : open(path, O_RDONLY);
        struct stat st;
        if (fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                r->map = m; r->map_len = (size_t)st.st_size; r->map_fd = fd; r->plain_map = true;
            }
        }
        if (!r->plain_map) {
            if (fd >= 0) close(fd);
            r->fp = fopen(path, "rb");                                  // pipes, empty files, AQC_READER_MMAP=0
        }
    }
    if (!r->gz && !r->fp && !r->inf && !r->pinf && !r->plain_map) { delete r; return AQC_ERR_INVALID; }
    if (r->gz || r->inf || r->pinf) r->async = new AsyncSource([r](uint8_t *dst, size_t cap) { return r->read_gz(dst, cap); }, kReadBlock, 3);
    r->raw = raw;
    if (raw) { *out = r; return 0; }
    r->slots.resize(slots);
    for (uint32_t i = 0; i < slots; i++) {
        Slot &s = r->slots[i];
        for (int c = 0; c < 4; c++) s.off[c] = (uint64_t *)malloc((batch_records + 1) * sizeof(uint64_t));
        s.off32 = (uint32_t *)malloc((batch_records + 1) * sizeof(uint32_t));
        if (!s.off[0] || !s.off[1] || !s.off[2] || !s.off[3] || !s.off32) { aqc_reader_close(r); return AQC_ERR_NOMEM; }
        r->free_.push_back(i);
    }
    r->th = std::thread([r] { r->run(); });
    *out = r;
    return 0;
}

int aqc_reader_next(aqc_reader *r, aqc_records *out) {
    if (!r || !out) return AQC_ERR_INVALID;
    std::unique_lock<std::mutex> lk(r->mu);
    r->cv.wait(lk, [&] { return !r->ready.empty() || r->finished; });
    memset(out, 0, sizeof *out);
    if (r->ready.empty()) { out->slot = UINT32_MAX; return r->err; }       // end of file (or the error, after the good batches)
    uint32_t id = r->ready.front(); r->ready.pop_front();
    const Slot &s = r->slots[id];
    out->n = s.n; out->first_index = s.first; out->slot = id; out->max_len = s.max_len;
    for (int c = 0; c < 4; c++) { out->bytes[c] = s.bytes[c]; out->off[c] = s.off[c]; }
    out->seq_off32 = s.off32;
    return 0;
}

int aqc_reader_release(aqc_reader *r, uint32_t slot) {
    if (!r || slot >= r->slots.size()) return AQC_ERR_INVALID;
    std::unique_lock<std::mutex> lk(r->mu);
    r->free_.push_back(slot);
    r->cv.notify_all();
    return 0;
}

// test / tool hook: gunzip a memory buffer with the reader's own decoder (all members); AQC_ERR_INVALID on a corrupt
// stream, AQC_ERR_NOMEM when out_cap is too small.
int aqc_gunzip_buffer(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t *out_len, char *err, uint64_t err_cap) {
    if (!in || !out_len) return AQC_ERR_INVALID;
    aqc::GzipInflater inf(in, (size_t)n);
    uint64_t w = 0;
    for (;;) {
        if (w == out_cap) {                                              // full: only fine if the stream ends here
            uint8_t probe;
            long g = inf.read(&probe, 1);
            if (g == 0) break;
            if (g < 0) { if (err && err_cap) snprintf(err, err_cap, "%s", inf.error().c_str()); return AQC_ERR_INVALID; }
            return AQC_ERR_NOMEM;
        }
        long g = inf.read(out + w, (size_t)(out_cap - w));
        if (g < 0) { if (err && err_cap) snprintf(err, err_cap, "%s", inf.error().c_str()); return AQC_ERR_INVALID; }
        if (g == 0) break;
        w += (uint64_t)g;
    }
    *out_len = w;
    return 0;
}

// the multi-threaded decoder on a memory buffer (no size threshold: for tests); stats: rounds, pieces, false starts
int aqc_gunzip_buffer_mt(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t *out_len, int threads,
                         uint64_t stats[3], char *err, uint64_t err_cap) {
    if (!in || !out_len) return AQC_ERR_INVALID;
    aqc::ParallelGunzip inf(in, (size_t)n, threads);
    uint64_t w = 0;
    int rc = 0;
    for (;;) {
        uint8_t probe;
        long g = w < out_cap ? inf.read(out + w, (size_t)(out_cap - w)) : inf.read(&probe, 1);
        if (g < 0) { if (err && err_cap) snprintf(err, err_cap, "%s", inf.error().c_str()); rc = AQC_ERR_INVALID; break; }
        if (g == 0) break;
        if (w >= out_cap) { rc = AQC_ERR_NOMEM; break; }
        w += (uint64_t)g;
    }
    if (stats) { stats[0] = inf.stats().rounds; stats[1] = inf.stats().pieces; stats[2] = inf.stats().false_starts; }
    *out_len = w;
    return rc;
}

const char *aqc_reader_error(const aqc_reader *r) { return r ? r->msg.c_str() : "null reader"; }

void aqc_reader_close(aqc_reader *r) {
    if (!r) return;
    {
        std::unique_lock<std::mutex> lk(r->mu);
        r->stop = true;
        r->cv.notify_all();
    }
    if (r->th.joinable()) r->th.join();
    for (Slot &s : r->slots) {
        for (int c = 0; c < 4; c++) { free(s.bytes[c]); free(s.off[c]); }
        free(s.off32);
    }
    delete r->async;                        // joins the decoder thread before its source goes away
    if (r->gz) gzclose(r->gz);
    if (r->fp) fclose(r->fp);
    delete r->inf;
    delete r->pinf;
    if (r->map) munmap(r->map, r->map_len);
    if (r->map_fd >= 0) close(r->map_fd);
    delete r;
}

}  // extern "C"
