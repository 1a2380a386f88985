// bamparallel.h — multi-threaded BAM ingest: BGZF blocks are inflated and BAM records are
// converted to rsq_read/rsq_block by a pool of worker threads; only the cheap framing of the record stream and the
// order-dependent scalar state of the reference loop (Decoder::observe) stay sequential.
//
// The reference decodes on the loop's own thread through SeqLib -> htslib sam_read1 (src/BamReader.cpp:12-27,
// SeqLib/src/BamReader.cpp:233-256); htslib's thread pool exists there (SeqLib/src/BamReader.cpp:176-183) but rnaseqc
// never enables it.  Here ingest scales with host cores so that the GPU hot path is fed at tens of M records/s.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "alnreader.h"
#include "rsq_b200.h"

namespace rsqh {

class ParallelBamDecoder {
    struct Job;                        // one ~4 MB piece of the decompressed stream (bamparallel.cpp)
public:
    // n_threads worker threads (>= 1) besides one file-reader thread; throws HostError(10) if the file cannot be opened
    ParallelBamDecoder(const std::string& path, int n_threads);
    ~ParallelBamDecoder();
    const AlnHeader& header() const { return hdr_; }
    // Starts the reader and the workers NOW: inflate, framing and conversion need only the command line's decode options and the
    // BAM header, so the CLI calls this before it parses the GTF and brings the device up — by the time the first batch is asked
    // for, up to max_jobs_ahead_ jobs (~1 MB of converted records each; RSQ_INGEST_AHEAD_JOBS) are waiting.  next_batch() starts them itself otherwise.  The
    // Decoder handed to next_batch() must have been built with the same options.
    void start(const DecodeOptions& opt);
    // Appends up to max_records converted records (with their blocks) to the cleared vectors, in file order, running
    // dec.observe() on each.  Returns false once the stream is exhausted and nothing was appended.
    bool next_batch(Decoder& dec, size_t max_records, std::vector<rsq_read>& reads, std::vector<rsq_block>& blocks);
    // The same stream as a list of pieces that stay where the workers wrote them (rsq_segment, include/rsq_b200.h): a job that goes
    // through its digest costs the caller's thread O(1) — no copy; records that need Decoder::observe one by one (first job, contig
    // changes, jobs cut by a batch boundary, the records that straddle two jobs) are copied, bound and observed into the batch's own
    // small arrays.  The jobs stay alive until recycle(); the contig ids of job pieces are bound by the packer (tids).
    struct Segments {
        std::vector<rsq_segment> segs;
        size_t n_reads = 0, n_blocks = 0;
        std::vector<std::shared_ptr<Job>> jobs;
        std::vector<rsq_read> loose_reads;            // the copied pieces, in stream order
        std::vector<rsq_block> loose_blocks;
        struct Loose { size_t seg, read0, n, block0, nb; };
        std::vector<Loose> loose;
    };
    bool next_batch_segments(Decoder& dec, size_t max_records, Segments& out);
    void recycle(Segments& s);
    // compressed bytes of the jobs handed out so far / size of the file: lets a caller extrapolate the record count
    uint64_t compressed_consumed() const { return compressed_consumed_; }
    uint64_t file_bytes() const { return file_bytes_; }
    // framing statistics of the decoder destroyed last (tests, diagnostics)
    static unsigned long long last_chain_hits();
    static unsigned long long last_chain_misses();
    static unsigned long long last_runs_summarised();          // jobs whose records went through Decoder::observe_run

private:
    void reader_loop();
    void worker_loop();
    void speculate(Job& j) const;                              // inflate stage: guessed record chain, converted
    bool frame_on_chain(Job& j, size_t begin, size_t end);     // frame stage: O(1) when the true start lies on that chain

    std::string path_;
    AlnHeader hdr_;
    uint64_t skip_bytes_ = 0;          // uncompressed bytes of the BAM header in front of the first record
    FILE* fp_ = nullptr;
    int n_threads_;
    std::thread reader_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_consumer_, cv_reader_;
    std::deque<std::shared_ptr<Job>> jobs_;        // in file order (ids are consecutive); front = oldest not yet consumed
    std::deque<std::shared_ptr<Job>> to_inflate_, to_convert_;   // work lists of the stages that any worker may take, oldest first
    std::vector<std::shared_ptr<Job>> pool_;       // consumed jobs whose buffers are reused
    std::vector<std::vector<uint8_t>> out_pool_;   // 4 MB inflate buffers: a job gives its buffer back once its records are converted
    std::vector<std::vector<uint8_t>> comp_pool_;  // ... and the buffer of its compressed bytes
    size_t n_unconverted_ = 0;                     // jobs that still hold an inflate buffer (or are about to take one)
    uint64_t next_job_id_ = 0;
    uint64_t next_frame_id_ = 0;       // framing is sequential: the next job id to frame
    std::vector<uint8_t> carry_;       // partial record at the end of the last framed job
    bool reader_done_ = false, stop_ = false;
    size_t max_jobs_in_flight_;        // ... of those: bounds the inflate buffers
    size_t max_jobs_ahead_;            // ... of all jobs not yet consumed: bounds the converted records waiting for the consumer
    // conversion parameters: the command line's decode options (start) and the header's number of references
    bool started_ = false;
    DecodeOptions conv_opt_;
    // consumer cursor
    std::shared_ptr<Job> cur_;
    size_t cur_read_ = 0;
    // busy seconds per stage, printed by the destructor when RSQ_INGEST_STATS is set (reader and consumer: one thread each;
    // frame: one worker at a time; inflate and convert: summed over the workers)
    uint64_t compressed_consumed_ = 0, file_bytes_ = 0;
    bool speculate_ = true;            // RSQ_INGEST_NO_SPECULATION=1: the sequential hop for every job (A/B, tests)
    bool summarise_ = true;            // RSQ_INGEST_NO_SUMMARY=1: Decoder::observe per record on the consumer thread (A/B, tests)
    uint64_t n_chain_hits_ = 0, n_chain_misses_ = 0, n_runs_summarised_ = 0;
    double t_reader_ = 0, t_frame_ = 0, t_inflate_ = 0, t_speculate_ = 0, t_convert_ = 0, t_consumer_ = 0, t_consumer_wait_ = 0;
};

}  // namespace rsqh
