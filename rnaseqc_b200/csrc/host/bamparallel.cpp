// bamparallel.cpp — multi-threaded BAM ingest (see bamparallel.h).
//
// Pipeline (all stages overlap):
//   reader thread : sequential file read, BGZF block framing, groups ~4 MB of blocks into a job
//   workers       : inflate a job (raw deflate per block)           -> job.out
//                   and SPECULATE while its bytes are still in this core's cache: guess where the first record of the job
//                   starts (the first offset from which several plausible records follow one another), hop to the end of
//                   the job along block_size fields and convert every record on that chain
//                   FRAME the next job in file order (sequential): the true start of the job's first fresh record follows
//                   from the partial record carried over from the previous job; if it lies on the speculated chain (hopping
//                   is deterministic, so the chain from there on IS the record stream) the job is done but for the one carried
//                   record — O(1) sequential work per job.  Otherwise (wrong guess, corrupt bytes, a record longer than a job,
//                   the BAM header) the sequential hop over block_size fields of the first version runs, followed by
//                   CONVERT of the framed job: every whole BAM record -> rsq_read + rsq_block (Decoder::convert_bam_bytes, pure)
//                   (the dependent hop through 4 MB another core had just written cost ~100 ns per record and capped ingest
//                   at ~10 M records/s whatever the number of workers)
//   consumer      : next_batch(): takes converted jobs in file order, binds the annotation's contig ids (Decoder::bind_contig),
//                   runs Decoder::observe (readLength, contig-drop state, unsorted warning: src/RNASeQC.cpp:275-278,346-356)
//                   and appends to the caller's batch
// Nothing before the consumer needs the annotation: start() runs all of it while the GTF is parsed and the device comes up.
// A job hands its 4 MB inflate buffer back as soon as its records are converted, so the jobs waiting for the consumer cost
// about 1 MB each and the run-ahead can be deep (max_jobs_ahead_: 3072 jobs, ~45 M records of 150 bp reads).
#include "bamparallel.h"
#include "fastinflate.h"

#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>

namespace rsqh {

namespace {
constexpr size_t kJobUncompressed = 4u << 20;   // target decompressed bytes per job
constexpr size_t kHeadroom = 256u << 10;        // room in front of a job's output for the previous job's partial record

using clk = std::chrono::steady_clock;
inline double since(clk::time_point t) { return std::chrono::duration<double>(clk::now() - t).count(); }

std::atomic<unsigned long long> g_last_hits{0}, g_last_misses{0}, g_last_summarised{0};

enum { ST_READ = 0, ST_INFLATING, ST_INFLATED, ST_FRAMING, ST_FRAMED, ST_CONVERTING, ST_CONVERTED };

// Reads one BGZF block from fp into `dst` (appended).  Returns false at clean EOF.
bool read_bgzf_block(FILE* fp, std::vector<uint8_t>& dst, uint32_t& payload_off, uint32_t& payload_len, uint32_t& isize) {
    uint8_t h[12];
    size_t r = std::fread(h, 1, 12, fp);
    if (r == 0) return false;
    if (r < 12 || h[0] != 0x1f || h[1] != 0x8b) throw HostError(-1, "corrupt BGZF block header");
    const unsigned xlen = h[10] | (h[11] << 8);
    uint8_t extra[256];
    std::vector<uint8_t> big;
    uint8_t* ex = extra;
    if (xlen > sizeof(extra)) { big.resize(xlen); ex = big.data(); }
    if (std::fread(ex, 1, xlen, fp) != xlen) throw HostError(-1, "truncated BGZF block");
    int bsize = -1;
    for (size_t i = 0; i + 4 <= xlen;) {
        const unsigned slen = ex[i + 2] | (ex[i + 3] << 8);
        if (ex[i] == 'B' && ex[i + 1] == 'C' && slen == 2) bsize = ex[i + 4] | (ex[i + 5] << 8);
        i += 4 + slen;
    }
    if (bsize < 0) throw HostError(-1, "BGZF block without BC subfield");
    const size_t cdata = (size_t)bsize + 1 - 12 - xlen - 8;
    const size_t at = dst.size();
    dst.resize(at + cdata + 8);
    if (std::fread(dst.data() + at, 1, cdata + 8, fp) != cdata + 8) throw HostError(-1, "truncated BGZF block");
    memcpy(&isize, dst.data() + at + cdata + 4, 4);
    payload_off = (uint32_t)at;
    payload_len = (uint32_t)cdata;
    return true;
}

// One raw-deflate stream per thread, reset between blocks (inflateInit2 allocates and clears ~40 KB of state each time).
struct ThreadInflater {
    z_stream zs;
    bool ready = false;
    ~ThreadInflater() { if (ready) inflateEnd(&zs); }
};

void inflate_block(const uint8_t* src, size_t n, uint8_t* dst, size_t isize) {
    if (!isize) return;
    // the whole-buffer decoder first (fastinflate.h); whatever it refuses goes to zlib, which inflates it or reports the error
    static const bool use_fast = !std::getenv("RSQ_INGEST_ZLIB");
    if (use_fast && fast_inflate(src, n, dst, isize)) return;
    thread_local ThreadInflater ti;
    if (!ti.ready) {
        memset(&ti.zs, 0, sizeof(ti.zs));
        if (inflateInit2(&ti.zs, -15) != Z_OK) throw HostError(-1, "zlib init failed");
        ti.ready = true;
    } else if (inflateReset(&ti.zs) != Z_OK) throw HostError(-1, "zlib init failed");
    ti.zs.next_in = const_cast<uint8_t*>(src);
    ti.zs.avail_in = (uInt)n;
    ti.zs.next_out = dst;
    ti.zs.avail_out = (uInt)isize;
    const int rc = inflate(&ti.zs, Z_FINISH);
    if (rc != Z_STREAM_END) throw HostError(-1, "BGZF inflate failed");
}

// One BAM record at base[p..): 1 = whole and plausible, 2 = incomplete tail (to be carried), 0 = not a record as far as the
// fixed fields can tell (SAM spec 4.2: field sizes must add up inside block_size, reference ids in range, NUL-terminated
// name, CIGAR operations 0-8).  Only used to pick and vet a speculated chain; a chain that fails is simply not used.
int plausible_record(const uint8_t* base, size_t p, size_t end, int32_t n_ref) {
    if (p + 4 > end) return 2;
    uint32_t bs; memcpy(&bs, base + p, 4);
    if (bs < 32 || bs > (1u << 28)) return 0;
    if (p + 36 > end) return 2;
    const uint8_t* b = base + p + 4;
    int32_t tid, mtid, l_seq; uint16_t n_cigar;
    memcpy(&tid, b, 4); memcpy(&mtid, b + 20, 4); memcpy(&l_seq, b + 16, 4); memcpy(&n_cigar, b + 12, 2);
    const size_t l_name = b[8];
    if (tid < -1 || tid >= n_ref || mtid < -1 || mtid >= n_ref || l_seq < 0 || l_name == 0) return 0;
    const size_t fixed = 32 + l_name + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
    if (fixed > bs) return 0;
    if (p + 4 + (size_t)bs > end) return 2;
    if (b[32 + l_name - 1] != 0) return 0;
    for (size_t k = 0; k < n_cigar; ++k) if ((b[32 + l_name + 4 * k] & 0xf) > 8) return 0;
    return 1;
}
}  // namespace

struct ParallelBamDecoder::Job {
    uint64_t id = 0;
    std::vector<uint8_t> comp;
    size_t comp_bytes = 0;             // comp.size() before the buffer went back to the pool (file-position estimates)
    std::vector<uint32_t> c_off, c_len, u_len;
    std::vector<uint8_t> out;          // [kHeadroom | inflated bytes]
    size_t data_len = 0;
    size_t rec_begin = 0, rec_end = 0; // whole records inside `out`
    std::vector<uint8_t> big;          // only when the carried partial record exceeds the head-room
    bool use_big = false;
    std::vector<rsq_read> reads;
    std::vector<rsq_block> rblocks;
    std::vector<int32_t> tids;
    // speculation (inflate stage): record starts of the guessed chain inside `out`, ascending; reads[k] belongs to chain[k]
    std::vector<uint32_t> chain;
    size_t chain_end = 0;              // first byte after the last whole record of the chain
    bool spec_ok = false;
    size_t first_read = 0;             // the consumer starts at reads[first_read]
    RunSummary summary;                // Decoder::summarise over reads[0..): lets the consumer skip the per-record observe()
    bool has_summary = false;
    // the record carried over from the previous job, converted by the framer (speculation hit only)
    bool has_head = false;
    rsq_read head_read;
    int32_t head_tid = 0;
    std::vector<rsq_block> head_blocks;
    std::string error;
    int error_code = 0;
    int state = ST_READ;
};

ParallelBamDecoder::ParallelBamDecoder(const std::string& path, int n_threads)
    : path_(path), n_threads_(std::max(1, n_threads)) {
    fp_ = std::fopen(path.c_str(), "rb");
    if (!fp_) throw HostError(10, "Unable to open BAM file: " + path);
    if (std::fseek(fp_, 0, SEEK_END) == 0) { const long sz = std::ftell(fp_); if (sz > 0) file_bytes_ = (uint64_t)sz; }
    std::rewind(fp_);
    // header: sequential inflate of the first block(s)
    {
        std::vector<uint8_t> text;
        auto need = [&](size_t n) {
            while (text.size() < n) {
                std::vector<uint8_t> comp;
                uint32_t off, len, isize;
                if (!read_bgzf_block(fp_, comp, off, len, isize)) throw HostError(-1, "truncated BAM header");
                const size_t at = text.size();
                text.resize(at + isize);
                inflate_block(comp.data() + off, len, text.data() + at, isize);
            }
        };
        need(12);
        if (memcmp(text.data(), "BAM\1", 4) != 0) throw HostError(-1, "not a BAM file: " + path);
        int32_t l_text; memcpy(&l_text, text.data() + 4, 4);
        size_t p = 8 + (size_t)l_text;
        need(p + 4);
        int32_t n_ref; memcpy(&n_ref, text.data() + p, 4);
        p += 4;
        for (int i = 0; i < n_ref; ++i) {
            need(p + 4);
            int32_t l_name; memcpy(&l_name, text.data() + p, 4);
            need(p + 4 + l_name + 4);
            std::string name((const char*)text.data() + p + 4, l_name > 0 ? l_name - 1 : 0);
            int32_t l_ref; memcpy(&l_ref, text.data() + p + 4 + l_name, 4);
            hdr_.ref_names.push_back(name);
            hdr_.ref_lens.push_back(l_ref);
            p += 8 + (size_t)l_name;
        }
        skip_bytes_ = p;
        std::fseek(fp_, 0, SEEK_SET);
    }
    max_jobs_in_flight_ = (size_t)n_threads_ * 3 + 4;
    max_jobs_ahead_ = 3072;                                        // ~3 GB of converted records at most
    if (const char* e = std::getenv("RSQ_INGEST_AHEAD_JOBS")) max_jobs_ahead_ = (size_t)std::max(1L, std::atol(e));
    max_jobs_ahead_ = std::max(max_jobs_ahead_, max_jobs_in_flight_);
    speculate_ = !std::getenv("RSQ_INGEST_NO_SPECULATION");
    summarise_ = !std::getenv("RSQ_INGEST_NO_SUMMARY");
}

ParallelBamDecoder::~ParallelBamDecoder() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_work_.notify_all();
    cv_reader_.notify_all();
    cv_consumer_.notify_all();
    if (reader_.joinable()) reader_.join();
    for (auto& t : workers_) if (t.joinable()) t.join();
    jobs_.clear(); to_inflate_.clear(); to_convert_.clear();
    if (fp_) std::fclose(fp_);
    g_last_hits = n_chain_hits_; g_last_misses = n_chain_misses_; g_last_summarised = n_runs_summarised_;
    if (std::getenv("RSQ_INGEST_STATS"))
        std::fprintf(stderr, "ingest busy seconds: reader %.3f, frame %.3f, inflate %.3f + convert on the speculated chain %.3f + convert after framing %.3f over %d workers, consumer %.3f (+ %.3f waiting); jobs framed on the speculated chain %llu, by the sequential hop %llu; jobs observed through their digest %llu\n",
                     t_reader_, t_frame_, t_inflate_, t_speculate_, t_convert_, n_threads_, t_consumer_, t_consumer_wait_,
                     (unsigned long long)n_chain_hits_, (unsigned long long)n_chain_misses_, (unsigned long long)n_runs_summarised_);
}

unsigned long long ParallelBamDecoder::last_chain_hits() { return g_last_hits; }
unsigned long long ParallelBamDecoder::last_chain_misses() { return g_last_misses; }
unsigned long long ParallelBamDecoder::last_runs_summarised() { return g_last_summarised; }

void ParallelBamDecoder::reader_loop() {
    try {
        bool eof = false;
        while (!eof) {
            std::shared_ptr<Job> job;
            {   // recycle a consumed job: its buffers are already mapped (fresh 4 MB vectors cost a page fault per page)
                std::lock_guard<std::mutex> lk(mu_);
                if (!pool_.empty()) { job = pool_.back(); pool_.pop_back(); }
                if (job && job->comp.capacity() == 0 && !comp_pool_.empty()) { job->comp = std::move(comp_pool_.back()); comp_pool_.pop_back(); }
            }
            if (!job) {
                job = std::make_shared<Job>();
                std::lock_guard<std::mutex> lk(mu_);
                if (!comp_pool_.empty()) { job->comp = std::move(comp_pool_.back()); comp_pool_.pop_back(); }
            }
            job->comp.clear(); job->c_off.clear(); job->c_len.clear(); job->u_len.clear();
            job->reads.clear(); job->rblocks.clear(); job->tids.clear(); job->big.clear();
            job->use_big = false; job->error.clear(); job->error_code = 0; job->state = ST_READ;
            job->chain.clear(); job->chain_end = 0; job->spec_ok = false; job->first_read = 0; job->has_head = false; job->head_blocks.clear(); job->has_summary = false;
            size_t total = 0;
            const auto t_job = clk::now();
            while (total < kJobUncompressed) {
                uint32_t off, len, isize;
                if (!read_bgzf_block(fp_, job->comp, off, len, isize)) { eof = true; break; }
                job->c_off.push_back(off); job->c_len.push_back(len); job->u_len.push_back(isize);
                total += isize;
            }
            if (job->c_off.empty()) break;
            job->data_len = total;
            t_reader_ += since(t_job);
            std::unique_lock<std::mutex> lk(mu_);
            cv_reader_.wait(lk, [&] { return stop_ || (n_unconverted_ < max_jobs_in_flight_ && jobs_.size() < max_jobs_ahead_); });
            if (stop_) return;
            job->id = next_job_id_++;
            jobs_.push_back(job);
            to_inflate_.push_back(job);
            ++n_unconverted_;
            lk.unlock();
            cv_work_.notify_one();
        }
    } catch (HostError& e) {
        auto job = std::make_shared<Job>();
        job->error = e.what(); job->error_code = e.exit_code; job->state = ST_CONVERTED;
        std::lock_guard<std::mutex> lk(mu_);
        job->id = next_job_id_++;
        jobs_.push_back(job);
    }
    {
        std::lock_guard<std::mutex> lk(mu_);
        reader_done_ = true;
    }
    cv_work_.notify_all();
    cv_consumer_.notify_all();
}

// Inflate stage, on the worker that has just written the job's bytes: pick the first offset from which four plausible records
// follow one another, hop to the end of the job, convert every record on the way.  Pure (no shared state).
void ParallelBamDecoder::speculate(Job& j) const {
    const uint8_t* base = j.out.data();
    const size_t begin = kHeadroom, end = kHeadroom + j.data_len;
    const int32_t n_ref = (int32_t)hdr_.ref_names.size();
    const size_t scan_end = std::min(end, begin + (64u << 10));
    for (size_t s = begin; s < scan_end; ++s) {
        size_t p = s;
        int k = 0, v = 1;
        for (; k < 4 && (v = plausible_record(base, p, end, n_ref)) == 1; ++k) { uint32_t bs; memcpy(&bs, base + p, 4); p += 4 + (size_t)bs; }
        if (v == 0 || k == 0) continue;
        // the chain from s
        j.chain.clear(); j.reads.clear(); j.rblocks.clear(); j.tids.clear();
        j.reads.reserve(j.data_len / 200 + 16);
        p = s;
        try {
            while ((v = plausible_record(base, p, end, n_ref)) == 1) {
                uint32_t bs; memcpy(&bs, base + p, 4);
                rsq_read r; int32_t tid;
                Decoder::convert_bam_record(conv_opt_, n_ref, base + p + 4, bs, r, j.rblocks, tid);
                j.chain.push_back((uint32_t)p);
                j.reads.push_back(r);
                j.tids.push_back(tid);
                p += 4 + (size_t)bs;
            }
        } catch (...) { v = 0; }
        if (v == 2) {
            j.chain_end = p; j.spec_ok = true;
            if (summarise_) { j.summary = Decoder::summarise(conv_opt_, n_ref, j.reads.data(), j.tids.data(), j.reads.size()); j.has_summary = true; }
        }
        else { j.chain.clear(); j.reads.clear(); j.rblocks.clear(); j.tids.clear(); }
        return;                                  // one guess per job: a second would cost a second conversion
    }
}

// Frame stage (sequential): does the job's first fresh record start on the speculated chain?  Then the chain from there on is
// the record stream; only the record carried over from the previous job is left to convert.  Returns false (nothing changed)
// when the sequential hop has to run instead.
bool ParallelBamDecoder::frame_on_chain(Job& j, size_t begin, size_t end) {
    uint8_t* base = j.out.data();
    const size_t c = carry_.size();
    size_t fresh = begin;
    uint32_t bs = 0;
    if (c) {
        if (c > begin || c + (end - begin) < 4) return false;
        uint8_t len4[4];
        for (size_t k = 0; k < 4; ++k) len4[k] = k < c ? carry_[k] : base[begin + (k - c)];
        memcpy(&bs, len4, 4);
        if (bs < 32 || 4 + (size_t)bs < c) return false;
        const size_t need = 4 + (size_t)bs - c;
        if (begin + need > end) return false;
        fresh = begin + need;
    }
    const auto it = std::lower_bound(j.chain.begin(), j.chain.end(), (uint32_t)fresh);
    if (it == j.chain.end() || *it != fresh) return false;
    if (c) {
        memcpy(base + begin - c, carry_.data(), c);
        j.head_blocks.clear();
        Decoder::convert_bam_record(conv_opt_, (int)hdr_.ref_names.size(), base + begin - c + 4, bs, j.head_read, j.head_blocks, j.head_tid);   // may throw: the job's error, as in CONVERT
        j.head_read.block_off = 0;
        j.has_head = true;
    }
    j.first_read = (size_t)(it - j.chain.begin());
    carry_.assign(base + j.chain_end, base + end);
    return true;
}

void ParallelBamDecoder::worker_loop() {
    // The workers fill every core; the thread that matters for the wall clock is the caller's (GTF parse, then the consumer and the
    // submits).  A few nice levels keep it on a core when there are as many workers as cores (Linux: per-thread priority by tid).
    if (!std::getenv("RSQ_INGEST_NO_NICE")) setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), 5);
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
        if (stop_) return;
        auto& q = jobs_;
        std::shared_ptr<Job> todo;
        int what = 0;   // 1 frame, 2 convert, 3 inflate
        // 1. framing is sequential: only the job whose id is next_frame_id_, and only one framer at a time (carry_ is its state);
        //    ids are consecutive in jobs_, so that job sits at a known place
        if (!q.empty() && next_frame_id_ >= q.front()->id && next_frame_id_ - q.front()->id < q.size()) {
            auto& j = q[(size_t)(next_frame_id_ - q.front()->id)];
            if (j->state == ST_INFLATED) { todo = j; what = 1; j->state = ST_FRAMING; }
        }
        if (!todo && !to_convert_.empty()) { todo = to_convert_.front(); to_convert_.pop_front(); what = 2; todo->state = ST_CONVERTING; }
        if (!todo && !to_inflate_.empty()) {
            todo = to_inflate_.front(); to_inflate_.pop_front(); what = 3; todo->state = ST_INFLATING;
            if (todo->out.capacity() == 0 && !out_pool_.empty()) { todo->out = std::move(out_pool_.back()); out_pool_.pop_back(); }
        }
        if (!todo) {
            if (reader_done_ && q.empty()) return;
            cv_work_.wait(lk);
            continue;
        }
        lk.unlock();
        const auto t_stage = clk::now();
        bool framed_by_chain = false;
        double dt_spec = 0.0;
        try {
            if (what == 3) {
                if (todo->out.size() < kHeadroom + todo->data_len + 16) todo->out.resize(kHeadroom + todo->data_len + 16);
                size_t o = kHeadroom;
                for (size_t b = 0; b < todo->c_off.size(); ++b) {
                    inflate_block(todo->comp.data() + todo->c_off[b], todo->c_len[b], todo->out.data() + o, todo->u_len[b]);
                    o += todo->u_len[b];
                }
                if (speculate_) { const auto t_spec = clk::now(); speculate(*todo); dt_spec = since(t_spec); }
            } else if (what == 1) {
                // frame: [carry | data] must be contiguous; then hop over whole records
                uint8_t* base = todo->out.data();
                size_t begin = kHeadroom, end = kHeadroom + todo->data_len;
                if (skip_bytes_) {                           // BAM header bytes in front of the first record
                    const size_t s = (size_t)std::min<uint64_t>(skip_bytes_, todo->data_len);
                    begin += s; skip_bytes_ -= s;
                }
                if (todo->spec_ok && !skip_bytes_ && frame_on_chain(*todo, begin, end)) {
                    framed_by_chain = true;
                } else {
                todo->reads.clear(); todo->rblocks.clear(); todo->tids.clear(); todo->first_read = 0; todo->has_head = false; todo->has_summary = false;
                if (!carry_.empty()) {
                    if (carry_.size() <= begin) { begin -= carry_.size(); memcpy(base + begin, carry_.data(), carry_.size()); }
                    else {
                        const size_t joined = carry_.size() + (end - begin);
                        todo->big.resize(joined + 16);                     // slack: the aux walk may look a few bytes past a corrupt record
                        memcpy(todo->big.data(), carry_.data(), carry_.size());
                        memcpy(todo->big.data() + carry_.size(), base + begin, end - begin);
                        todo->use_big = true; base = todo->big.data(); begin = 0; end = joined;
                    }
                    carry_.clear();
                }
                size_t p = begin;
                while (p + 4 <= end) {
                    uint32_t bs; memcpy(&bs, base + p, 4);
                    if (bs < 32) throw HostError(-1, "truncated BAM record");
                    if (p + 4 + (size_t)bs > end) break;
                    p += 4 + (size_t)bs;
                }
                todo->rec_begin = begin; todo->rec_end = p;
                carry_.assign(base + p, base + end);
                }
            } else {
                const uint8_t* base = todo->use_big ? todo->big.data() : todo->out.data();
                size_t p = todo->rec_begin;
                todo->reads.reserve((todo->rec_end - todo->rec_begin) / 200 + 16);
                while (p < todo->rec_end) {
                    uint32_t bs; memcpy(&bs, base + p, 4);
                    rsq_read r; int32_t tid;
                    Decoder::convert_bam_record(conv_opt_, (int)hdr_.ref_names.size(), base + p + 4, bs, r, todo->rblocks, tid);
                    todo->reads.push_back(r);
                    todo->tids.push_back(tid);
                    p += 4 + (size_t)bs;
                }
                if (summarise_) { todo->summary = Decoder::summarise(conv_opt_, (int)hdr_.ref_names.size(), todo->reads.data(), todo->tids.data(), todo->reads.size()); todo->has_summary = true; }
            }
        } catch (HostError& e) {
            todo->error = e.what(); todo->error_code = e.exit_code ? e.exit_code : -1;
        } catch (std::exception& e) {
            todo->error = e.what(); todo->error_code = -1;
        }
        const double dt_stage = since(t_stage);
        lk.lock();
        (what == 3 ? t_inflate_ : what == 1 ? t_frame_ : t_convert_) += dt_stage - dt_spec;
        t_speculate_ += dt_spec;
        if (!todo->error.empty()) {
            // a failed job surfaces at the consumer when its turn comes; framing moves past it so earlier records still flow
            if (todo->id == next_frame_id_) ++next_frame_id_;
            todo->state = ST_CONVERTED;
        } else if (what == 3) todo->state = ST_INFLATED;
        else if (what == 1) {
            ++next_frame_id_;
            if (framed_by_chain) { todo->state = ST_CONVERTED; ++n_chain_hits_; }
            else { todo->state = ST_FRAMED; ++n_chain_misses_; to_convert_.push_back(todo); }
        }
        else todo->state = ST_CONVERTED;
        if (todo->state == ST_CONVERTED) {
            // the records are out of the inflated bytes: the 4 MB buffer goes to the next job, and the reader may read on
            if (todo->out.capacity()) { out_pool_.push_back(std::move(todo->out)); todo->out = std::vector<uint8_t>(); }
            todo->big = std::vector<uint8_t>(); todo->use_big = false;
            // ... and neither the compressed bytes nor the chain are looked at again: a waiting job keeps its records only
            todo->comp_bytes = todo->comp.size();
            if (todo->comp.capacity()) { comp_pool_.push_back(std::move(todo->comp)); todo->comp = std::vector<uint8_t>(); }
            todo->chain = std::vector<uint32_t>();
            --n_unconverted_;
            cv_consumer_.notify_all();
            cv_reader_.notify_one();
        }
        cv_work_.notify_all();
    }
}

void ParallelBamDecoder::start(const DecodeOptions& opt) {
    std::unique_lock<std::mutex> lk(mu_);
    if (started_) return;
    started_ = true;
    conv_opt_ = opt;
    reader_ = std::thread(&ParallelBamDecoder::reader_loop, this);
    for (int t = 0; t < n_threads_; ++t) workers_.emplace_back(&ParallelBamDecoder::worker_loop, this);
}

bool ParallelBamDecoder::next_batch(Decoder& dec, size_t max_records, std::vector<rsq_read>& reads, std::vector<rsq_block>& blocks) {
    reads.clear();
    blocks.clear();
    start(dec.options());
    while (reads.size() < max_records) {
        std::shared_ptr<Job> job = cur_;
        if (!job) {
            std::unique_lock<std::mutex> lk(mu_);
            auto& q = jobs_;
            const auto t_wait = clk::now();
            cv_consumer_.wait(lk, [&] { return stop_ || (!q.empty() && q.front()->state == ST_CONVERTED) || (reader_done_ && q.empty()); });
            t_consumer_wait_ += since(t_wait);
            if (stop_ || q.empty()) break;
            job = q.front();
            q.pop_front();
            lk.unlock();
            cv_reader_.notify_one();
            cv_work_.notify_all();
            if (!job->error.empty()) throw HostError(job->error_code, job->error);
            cur_ = job;
            cur_read_ = job->first_read;
            if (job->has_head) {                 // the record that straddles the previous job's end comes first
                rsq_read r = job->head_read;
                r.block_off = (uint32_t)blocks.size();
                blocks.insert(blocks.end(), job->head_blocks.begin(), job->head_blocks.end());
                dec.bind_contig(r, job->head_tid);
                reads.push_back(r);
                dec.observe(reads.back(), job->head_tid);
                job->has_head = false;
                if (reads.size() >= max_records) continue;
            }
        }
        const size_t n = std::min(job->reads.size() - cur_read_, max_records - reads.size());
        const auto t_copy = clk::now();
        if (n) {
            // bulk append: the records of a job and their blocks are contiguous and in file order
            const rsq_read* first = job->reads.data() + cur_read_;
            const uint32_t b_lo = first->block_off;
            const uint32_t b_hi = first[n - 1].block_off + first[n - 1].n_blocks;
            const uint32_t delta = (uint32_t)blocks.size() - b_lo;
            const size_t at = reads.size();
            reads.insert(reads.end(), first, first + n);
            blocks.insert(blocks.end(), job->rblocks.data() + b_lo, job->rblocks.data() + b_hi);
            const int32_t* tids = job->tids.data() + cur_read_;
            rsq_read* out = reads.data() + at;
            // a whole job whose digest shows that no record moves the order-dependent state (but for the last position): O(1)
            if (job->has_summary && cur_read_ == 0 && n == job->reads.size() && dec.observe_run(job->summary)) {
                ++n_runs_summarised_;
                for (size_t k = 0; k < n; ++k) { out[k].block_off += delta; dec.bind_contig(out[k], tids[k]); }
            } else {
                for (size_t k = 0; k < n; ++k) {
                    out[k].block_off += delta;
                    dec.bind_contig(out[k], tids[k]);
                    dec.observe(out[k], tids[k]);
                }
            }
        }
        t_consumer_ += since(t_copy);
        cur_read_ += n;
        if (cur_read_ >= job->reads.size()) {
            compressed_consumed_ += cur_->comp_bytes;
            { std::lock_guard<std::mutex> lk(mu_); pool_.push_back(cur_); }
            cur_.reset(); cur_read_ = 0;
        }
    }
    return !reads.empty();
}

// next_batch without the copy (bamparallel.h).  Same order of Decoder calls as next_batch: the stream state cannot tell the two apart.
bool ParallelBamDecoder::next_batch_segments(Decoder& dec, size_t max_records, Segments& out) {
    out.segs.clear(); out.jobs.clear(); out.loose_reads.clear(); out.loose_blocks.clear(); out.loose.clear();
    out.n_reads = 0; out.n_blocks = 0;
    start(dec.options());
    auto loose_piece = [&]() -> Segments::Loose& {          // the copied piece at the end of the batch so far (a new one if a job piece is last)
        if (out.loose.empty() || out.loose.back().seg + 1 != out.segs.size()) {
            out.loose.push_back(Segments::Loose{out.segs.size(), out.loose_reads.size(), 0, out.loose_blocks.size(), 0});
            out.segs.push_back(rsq_segment{nullptr, nullptr, 0, nullptr, 0});
        }
        return out.loose.back();
    };
    while (out.n_reads < max_records) {
        std::shared_ptr<Job> job = cur_;
        if (!job) {
            std::unique_lock<std::mutex> lk(mu_);
            auto& q = jobs_;
            const auto t_wait = clk::now();
            cv_consumer_.wait(lk, [&] { return stop_ || (!q.empty() && q.front()->state == ST_CONVERTED) || (reader_done_ && q.empty()); });
            t_consumer_wait_ += since(t_wait);
            if (stop_ || q.empty()) break;
            job = q.front();
            q.pop_front();
            lk.unlock();
            cv_reader_.notify_one();
            cv_work_.notify_all();
            if (!job->error.empty()) throw HostError(job->error_code, job->error);
            cur_ = job;
            cur_read_ = job->first_read;
            if (job->has_head) {                 // the record that straddles the previous job's end comes first
                Segments::Loose& L = loose_piece();
                rsq_read r = job->head_read;
                r.block_off = (uint32_t)(out.loose_blocks.size() - L.block0);
                out.loose_blocks.insert(out.loose_blocks.end(), job->head_blocks.begin(), job->head_blocks.end());
                dec.bind_contig(r, job->head_tid);
                out.loose_reads.push_back(r);
                dec.observe(out.loose_reads.back(), job->head_tid);
                L.n += 1; L.nb += job->head_blocks.size();
                out.n_reads += 1; out.n_blocks += job->head_blocks.size();
                job->has_head = false;
                if (out.n_reads >= max_records) continue;
            }
        }
        const size_t n = std::min(job->reads.size() - cur_read_, max_records - out.n_reads);
        const auto t_copy = clk::now();
        bool by_reference = false;
        if (n) {
            const rsq_read* first = job->reads.data() + cur_read_;
            const uint32_t b_lo = first->block_off;
            const uint32_t b_hi = first[n - 1].block_off + first[n - 1].n_blocks;
            if (job->has_summary && cur_read_ == 0 && n == job->reads.size() && dec.observe_run(job->summary)) {
                ++n_runs_summarised_;
                out.segs.push_back(rsq_segment{first, job->tids.data(), n, job->rblocks.data(), job->rblocks.size()});
                out.jobs.push_back(job);
                by_reference = true;
            } else {
                Segments::Loose& L = loose_piece();
                const size_t at = out.loose_reads.size();
                const uint32_t delta = (uint32_t)(out.loose_blocks.size() - L.block0) - b_lo;
                out.loose_reads.insert(out.loose_reads.end(), first, first + n);
                out.loose_blocks.insert(out.loose_blocks.end(), job->rblocks.data() + b_lo, job->rblocks.data() + b_hi);
                const int32_t* tids = job->tids.data() + cur_read_;
                rsq_read* o = out.loose_reads.data() + at;
                for (size_t k = 0; k < n; ++k) {
                    o[k].block_off += delta;
                    dec.bind_contig(o[k], tids[k]);
                    dec.observe(o[k], tids[k]);
                }
                L.n += n; L.nb += b_hi - b_lo;
            }
            out.n_reads += n; out.n_blocks += b_hi - b_lo;
        }
        t_consumer_ += since(t_copy);
        cur_read_ += n;
        if (cur_read_ >= job->reads.size()) {
            compressed_consumed_ += cur_->comp_bytes;
            if (!by_reference) { std::lock_guard<std::mutex> lk(mu_); pool_.push_back(cur_); }   // a referenced job goes back in recycle()
            cur_.reset(); cur_read_ = 0;
        }
    }
    for (const Segments::Loose& L : out.loose) {          // the copied pieces have stopped moving
        rsq_segment& S = out.segs[L.seg];
        S.reads = out.loose_reads.data() + L.read0; S.tids = nullptr; S.n_reads = L.n;
        S.blocks = out.loose_blocks.data() + L.block0; S.n_blocks = L.nb;
    }
    return out.n_reads != 0;
}

void ParallelBamDecoder::recycle(Segments& s) {
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto& j : s.jobs) pool_.push_back(std::move(j));
    }
    s.jobs.clear(); s.segs.clear();
}

}  // namespace rsqh
