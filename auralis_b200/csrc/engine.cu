// Engine: weights, paged-KV slots, continuous-batching scheduler thread, vocoder driver and the C ABI
// (include/xtts_b200.h).  Host orchestration that stands in for XTTSv2Engine + vLLM's engine loop
// (XTTSv2.py:690-814); every tensor op below it is one of the kernels in this directory.
#include "../../include/xtts_b200.h"
#include "kernels.h"
#include "cond.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

namespace xtts {

static thread_local std::string t_last_error;
static std::string g_last_error;
static std::mutex g_err_mu;

static void set_error(const std::string& s) {
    t_last_error = s;
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_last_error = s;
}

static double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------------
template <typename T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
    }
    void zero(cudaStream_t st) { if (n) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), st)); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    void upload(const T* h, size_t count, cudaStream_t st, size_t off = 0) {
        CUDA_CHECK(cudaMemcpyAsync(p + off, h, count * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void download(T* h, size_t count, cudaStream_t st, size_t off = 0) const {
        CUDA_CHECK(cudaMemcpyAsync(h, p + off, count * sizeof(T), cudaMemcpyDeviceToHost, st));
    }
};

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
    size_t numel() const { size_t n = 1; for (auto s : shape) n *= (size_t)s; return n; }
};

// weight matrix usable as the B operand of either GEMM path
struct Linear {
    DBuf<float> w32;              // [N,K] fp32   (precision fp32)
    DBuf<__nv_bfloat16> w16;      // [N,K] bf16   (precision bf16)
    DBuf<float> b;                // [N]
    int N = 0, K = 0;
};

struct ConvW {
    DBuf<float> wt;               // [Cin][K][Cout]      (CUDA-core path)
    DBuf<__half> blob;            // tcgen05 tile blob   (fast mode, Conv1d only)
    ConvTcPlan plan{};
    bool tc = false;
    DBuf<float> b;                // [Cout]
    int Cin = 0, Cout = 0, K = 0;
};

struct Sequence {
    uint64_t id = 0;
    std::vector<int32_t> text_ids;
    int speaker = 0;
    xtts_sampling sp{};
    int slot = -1;
    int n_prompt = 0;
    int steps = 0;                // decode steps issued so far (host-side mirror of n_gen - 1)
    std::vector<int> pages;
    double t_submit = 0, t_first = 0, t_done = 0;
    // results
    int status = 0;
    std::vector<int32_t> tokens;
    int n_samples = 0;
    float* wav_host = nullptr;    // pinned
    size_t wav_cap = 0;
    float* wav_dev = nullptr;     // kept when d2h_wav == 0 (pooled device buffer)
    size_t wav_dev_cap = 0;
    float* lat_dev = nullptr;     // [n_tokens, H] copy so the slot can be reused (pooled device buffer)
    size_t lat_dev_cap = 0;
    // first-audio early emit (xtts_sampling.early_tokens; 0 = off)
    int early_tokens = 0;         // audio of this many leading tokens goes out as a partial result
    bool early_done = false;
    int early_samples = 0;        // samples already delivered by the partial result
};

// a partial (first-audio) piece lives in done_map under its sequence id with this bit set; user ids must not use it
static constexpr uint64_t kPartialBit = 1ull << 63;
// latent frames decoded beyond `early_tokens` before the prefix is vocoded: the vocoder's receptive field reaches
// ~3 latent frames ahead (conv_pre + the k = 11, d = 5 resblock of the first stage; measured with the oracle), 6 leaves margin
static constexpr int kEarlyLookahead = 6;

class Engine {
public:
    explicit Engine(const xtts_config& c);
    ~Engine();

    void load_weight(const char* name, const float* data, const int64_t* shape, int ndim);
    void finalize_weights();
    void set_speaker(int slot, const float* cond, const float* g);
    void get_speaker(int slot, float* cond, float* g);
    void condition(int slot, const float* w22, int64_t n22, const float* w16, int64_t n16, int cond_len, int chunk_len);
    void submit(uint64_t id, const int32_t* text, int n_text, int speaker, const xtts_sampling& sp);
    int poll(xtts_result* out, int timeout_ms);
    void fetch(uint64_t id, int32_t* tokens, float* wav, float* latents);
    void set_option(const std::string& k, int64_t v);
    void get_stats(xtts_stats* s);
    void sync_idle();
    void kernel_profile(xtts_kernel_profile* out);
    void device_timer(int op, double* ms);

    void vocode_sync(const float* latents, int T, int speaker, float* wav, int* n_out, const char* stage,
                     float* stage_out, int64_t stage_cap);
    void gpt_prefill_sync(const int32_t* text, int n_text, int speaker, const int32_t* audio, int n_audio,
                          float* hidden_out, float* logits_out, float* latents_out);
    void gpt_teacher_forced_sync(const int32_t* text, int n_text, int speaker, const int32_t* forced, int n,
                                 const xtts_sampling& sp, float* logits_out, float* latents_out, int32_t* sampled_out);
    void debug_gemm(int mode, const float* A, const float* W, const float* bias, const float* resid, float* out, int M,
                    int N, int K, int gelu, int iters, float* ms);
    void debug_sample(const float* logits, const uint8_t* seen, int B, int V, const xtts_sampling& sp, int step,
                      int32_t* out);

private:
    // ---- geometry
    xtts_config cfg;
    int H, L, NH, FF, V, Vpad, B, NSLOT, CAP, MAXP, max_pages, SEENW, S;
    int prefill_rows_cap;
    bool bf16;
    cudaStream_t st = nullptr;
    static constexpr int kMaxMicro = 4;
    cudaStream_t st_mb[kMaxMicro] = {nullptr, nullptr, nullptr, nullptr};   // [0] == st; decode micro-batch branches
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;       // xtts_device_timer
    bool timer_armed = false;
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxMicro] = {nullptr, nullptr, nullptr, nullptr};

    // ---- weights
    std::map<std::string, HostTensor> raw;
    std::atomic<bool> finalized{false};
    DBuf<float> text_emb, text_pos, wte, wpe;
    struct Layer { DBuf<float> ln1w, ln1b, ln2w, ln2b; Linear qkv, o, fc, proj; };
    std::vector<std::unique_ptr<Layer>> layers;
    DBuf<float> lnfw, lnfb, fnw, fnb;
    Linear mel_head;
    ConvW conv_pre;
    std::vector<std::unique_ptr<ConvW>> ups;
    struct RB { std::unique_ptr<ConvW> c1[4], c2[4]; };
    std::vector<std::unique_ptr<RB>> rbs;
    DBuf<float> conv_post_w;
    int post_cin = 0;
    struct CondLin { DBuf<float> w, b; int rows = 0; };
    CondLin cond_layer;
    std::vector<std::unique_ptr<CondLin>> conds;
    uint64_t weight_bytes = 0;
    std::unique_ptr<Conditioner> conditioner;

    // ---- speakers
    DBuf<float> spk_cond, spk_g, spk_cbias;
    int cbias_stride = 0;
    std::vector<int> cbias_off;
    std::vector<char> spk_valid;

    // ---- slot state (device)
    DBuf<int> d_last_tok, d_n_gen, d_ctx_len, d_finished, d_tokens, d_sampled, d_forced, d_top_k, d_max_tokens,
        d_stop, d_seq_seed, d_block_tables, d_active, d_rowidx, d_row_slot, d_row_pos, d_lat_pos;
    DBuf<unsigned> d_seen;
    DBuf<float> d_temp, d_top_p, d_pen;
    DBuf<unsigned long long> d_seed;
    DBuf<float> d_latents;
    DBuf<RowDesc> d_rows;
    DBuf<AttnSeq> d_attnseq;
    bool use_forced = false;
    // KV pools, one pair per layer
    std::vector<std::unique_ptr<DBuf<float>>> k32, v32;
    std::vector<std::unique_ptr<DBuf<__nv_bfloat16>>> k16, v16;
    std::vector<int> free_pages;
    // GPT workspace
    DBuf<float> wX, wQKV, wLOG, wPART;
    DBuf<float> wXn32, wATT32, wFF32, wY32;
    DBuf<__nv_bfloat16> wXn16, wATT16, wFF16, wY16;
    // pinned staging
    int* h_finished = nullptr;
    // vocoder workspace
    DBuf<float> vz, vpre, vb[5], vwav, vlat, vcb;
    DBuf<__half> vz16, va16[5];   // fp16 operand atoms of the tensor-core vocoder path
    bool tc_vocoder_ready = false;
    int voc_max_T = 0, VB = 1;
    std::vector<int> stage_ch;

    // ---- scheduler
    // Two locks so that submit/poll never wait behind a GPU step:
    //   q_mu   — pending / done queues, inflight counter (short critical sections)
    //   mu     — all GPU state and work issue: held by the scheduler thread for one iteration at a time and by
    //            the synchronous entry points (weights, speakers, debug calls)
    std::mutex mu, q_mu, pin_mu;
    std::condition_variable cv_work, cv_done;
    std::deque<std::shared_ptr<Sequence>> pending;      // q_mu
    std::deque<std::shared_ptr<Sequence>> waiting;      // scheduler thread only: accepted, not yet admitted
    std::vector<std::shared_ptr<Sequence>> running;     // index = position in active list
    std::vector<int> free_slots;
    std::deque<std::shared_ptr<Sequence>> done_q;
    std::unordered_map<uint64_t, std::shared_ptr<Sequence>> done_map;
    std::thread worker;
    std::atomic<bool> stop{false};
    int inflight = 0;
    bool d2h_wav = true;
    bool use_splitk = true;       // option "splitk"
    bool use_pdl = true;          // option "pdl" (programmatic dependent launch along the decode chain)
    bool use_graphs = true;       // option "cuda_graphs"
    std::atomic<bool> hold_admission{false};   // option "hold_admission": queue submissions, admit nothing (batch submit)
    std::map<int, std::pair<cudaGraphExec_t, KernelProfiler::GraphRecs*>> prof_graphs;   // decode graphs with event-record nodes (option "profile")
    KernelProfiler::GraphRecs* last_prof = nullptr;     // recs of the profiled graph launched by the current step
    void drop_graphs();
    bool use_chain = false;       // option "decode_chain": fused persistent per-layer GEMM/LayerNorm chain kernel (measured slower)
    DBuf<unsigned> d_chain_sync;  // device-wide barrier words of the chain kernel
    int n_micro = 2;              // option "microbatches": decode rows are split into this many concurrent branches
    int micro_min_rows = 48;      // option "microbatch_min_rows": below this many active rows the step stays single-branch
    int eager_steps_done = 0;
    std::map<int, cudaGraphExec_t> decode_graphs;
    std::map<int, unsigned long long> graph_kernels;
    bool use_tc_vocoder = true;   // fast mode only; option "tc_vocoder" switches back to the fp32 CUDA-core convs
    std::vector<std::pair<float*, size_t>> pinned_pool;
    // stats
    double decode_ctx_sum = 0;
    uint64_t st_decode_steps = 0, st_prefill_rows = 0, st_tokens = 0, st_samples = 0;
    double st_gpt_ms = 0, st_voc_ms = 0, st_cond_ms = 0;
    unsigned long long launch_base = 0;

    // ---- helpers
    const HostTensor& need(const std::string& name) const;
    void up(DBuf<float>& d, const std::vector<float>& h) { d.alloc(h.size()); d.upload(h.data(), h.size(), st); weight_bytes += h.size() * 4; }
    void make_linear(Linear& lin, const std::string& wname, const std::string& bname, bool conv1d_layout, int pad_n = 0);
    void make_conv(ConvW& c, const std::string& prefix, bool transposed, bool has_bias);
    void conv1d(const ConvW& c, const float* x, const float* cbias, const float* resid, float* out, int L, int dil,
                float in_scale, float slope, int mode, int nb);
    void conv1d_tc(const ConvW& c, const __half* a16, const float* cbias, const float* resid, float* out32, __half* out16,
                   int L, int lpad, int dil, int mode, int nb, float scale16 = 1.0f);
    void run_vocoder_tc(const float* lat_dev, int T, int nb, float* wav_dev_out, int* n_out, const char* stage,
                        float* stage_out, int64_t stage_cap);
    std::vector<float> folded(const std::string& prefix) const;
    GptTables tables() const {
        GptTables t; t.text_emb = text_emb.p; t.text_pos = text_pos.p; t.wte = wte.p; t.wpe = wpe.p;
        t.spk_cond = spk_cond.p; t.H = H; t.n_cond = cfg.n_cond_latents; return t;
    }
    SampleState sample_state() const;
    void finish_speaker(int slot);
    void gemm(const void* A, const Linear& lin, const float* resid, void* out, int M, int flags, bool pdl = false);
    void layers_forward(int M, bool prefill, int nseq, int max_nq);
    void head_and_sample(int M, const int* row_index, const int* slots_dev, const int* lat_pos, int advance_ctx, bool do_sample,
                         bool pdl_first = true);
    void decode_layers_rows(int r0, int Mi, cudaStream_t s, bool pdl_first, double ctx_sum);
    void decode_layers_chain(int M);
    void init_slot(Sequence& s, const int32_t* forced, int n_forced);
    void release_slot(Sequence& s);
    int build_prefill(const std::vector<Sequence*>& seqs, const std::vector<std::vector<int32_t>>& audio,
                      std::vector<int>& last_rows, int& max_nq);
    void prefill(const std::vector<Sequence*>& seqs);
    void decode_step(const std::vector<int>& active);
    void run_vocoder(const float* lat_dev, int T, const int* speakers, int nb, float* wav_dev_out, int* n_out,
                     const char* stage, float* stage_out, int64_t stage_cap);
    void finish_group(std::vector<std::shared_ptr<Sequence>>& grp);
    void emit_early(std::vector<std::shared_ptr<Sequence>>& grp);
    int samples_for(int T) const;
    void finish_sequence(std::shared_ptr<Sequence> s);
    void retire(std::shared_ptr<Sequence> s);
    float* pinned_get(size_t n, size_t* cap);
    void pinned_put(float* p, size_t cap);
    float* dev_get(size_t n, size_t* cap);           // device buffer pool (no cudaMalloc/cudaFree per chunk)
    void dev_put(float* p, size_t cap);
    std::vector<std::pair<float*, size_t>> dev_pool;
    void loop();
    void require_finalized() const { if (!finalized) throw std::runtime_error("weights not finalized: call xtts_finalize_weights first"); }
};

// ================================================================================================
// construction
// ================================================================================================
Engine::Engine(const xtts_config& c) : cfg(c) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        throw std::runtime_error("libxtts_b200: no CUDA device visible — this library has no CPU fallback");
    }
    if (c.device < 0 || c.device >= ndev) throw std::runtime_error("invalid CUDA device ordinal");
    CUDA_CHECK(cudaSetDevice(c.device));
    cudaDeviceProp prop{};
    CUDA_CHECK(cudaGetDeviceProperties(&prop, c.device));
    if (prop.major != 10) {
        throw std::runtime_error(std::string("libxtts_b200 is built for sm_100a only; found ") + prop.name + " (sm_" +
                                 std::to_string(prop.major) + std::to_string(prop.minor) + ")");
    }
    H = c.hidden; L = c.layers; NH = c.heads; FF = c.ff; V = c.n_audio_tokens;
    if (H != NH * kHeadDim) throw std::runtime_error("hidden must equal heads * 64");
    if (H % 64 != 0 || FF % 64 != 0) throw std::runtime_error("hidden and ff must be multiples of 64");
    if (V > 2048) throw std::runtime_error("n_audio_tokens > 2048 unsupported");
    Vpad = ceil_div(V, 32) * 32;
    B = c.max_batch; NSLOT = B + 1;     // last slot is reserved for the synchronous debug entry points
    CAP = c.max_audio_tokens;
    MAXP = c.n_cond_latents + (c.max_text_tokens + 2) + 1;
    max_pages = ceil_div(MAXP + CAP, kPageTokens);
    SEENW = ceil_div(V, 32);
    S = std::max(1, c.max_speakers);
    bf16 = c.precision == XTTS_PRECISION_BF16;
    if (bf16) { std::string err; if (!gemm_tc_init(&err)) throw std::runtime_error(err); }
    CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    st_mb[0] = st;
    for (int i = 1; i < kMaxMicro; ++i) CUDA_CHECK(cudaStreamCreateWithFlags(&st_mb[i], cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    for (int i = 1; i < kMaxMicro; ++i) CUDA_CHECK(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming));

    // speakers
    spk_cond.alloc((size_t)S * c.n_cond_latents * H);
    spk_g.alloc((size_t)S * c.d_vector);
    spk_valid.assign(S, 0);
    cbias_off.clear();
    int off = 0;
    cbias_off.push_back(off); off += c.voc_init_ch;
    stage_ch.clear();
    for (int i = 0; i < c.voc_n_up; ++i) {
        const int ch = c.voc_init_ch >> (i + 1);
        stage_ch.push_back(ch);
        cbias_off.push_back(off); off += ch;
    }
    cbias_stride = off;
    spk_cbias.alloc((size_t)S * cbias_stride);

    // slot state
    d_last_tok.alloc(NSLOT); d_n_gen.alloc(NSLOT); d_ctx_len.alloc(NSLOT); d_finished.alloc(NSLOT);
    d_tokens.alloc((size_t)NSLOT * CAP); d_sampled.alloc((size_t)NSLOT * CAP); d_forced.alloc((size_t)NSLOT * CAP);
    d_top_k.alloc(NSLOT); d_max_tokens.alloc(NSLOT); d_stop.alloc(NSLOT); d_seq_seed.alloc(NSLOT);
    d_temp.alloc(NSLOT); d_top_p.alloc(NSLOT); d_pen.alloc(NSLOT); d_seed.alloc(NSLOT);
    d_seen.alloc((size_t)NSLOT * SEENW);
    d_block_tables.alloc((size_t)NSLOT * max_pages);
    d_active.alloc(NSLOT);
    d_latents.alloc((size_t)NSLOT * CAP * H);
    d_finished.zero(st); d_n_gen.zero(st); d_ctx_len.zero(st); d_last_tok.zero(st);
    CUDA_CHECK(cudaMallocHost(&h_finished, NSLOT * sizeof(int)));

    // prefill row budget: whole prompts of up to 8 sequences at the maximum prompt length, >= one debug pass
    prefill_rows_cap = std::max(8 * MAXP, MAXP + CAP);
    d_rows.alloc(prefill_rows_cap); d_row_slot.alloc(prefill_rows_cap); d_row_pos.alloc(prefill_rows_cap);
    d_rowidx.alloc(prefill_rows_cap); d_lat_pos.alloc(prefill_rows_cap);
    d_attnseq.alloc(NSLOT);
    const size_t Mmax = (size_t)std::max(prefill_rows_cap, NSLOT);
    wX.alloc(Mmax * H); wQKV.alloc(Mmax * 3 * H); wLOG.alloc((size_t)std::max(NSLOT, CAP + 1) * Vpad);
    if (bf16) wPART.alloc((size_t)8 * NSLOT * H);
    d_chain_sync.alloc(64); d_chain_sync.zero(st);
    if (bf16) { wXn16.alloc(Mmax * H); wATT16.alloc(Mmax * H); wFF16.alloc(Mmax * FF); wY16.alloc((size_t)std::max(NSLOT, CAP + 1) * H); }
    else { wXn32.alloc(Mmax * H); wATT32.alloc(Mmax * H); wFF32.alloc(Mmax * FF); wY32.alloc((size_t)std::max(NSLOT, CAP + 1) * H); }

    // KV pools
    const int total_pages = NSLOT * max_pages;
    const size_t page_elems = (size_t)NH * kPageTokens * kHeadDim;
    for (int l = 0; l < L; ++l) {
        if (bf16) {
            k16.emplace_back(new DBuf<__nv_bfloat16>()); v16.emplace_back(new DBuf<__nv_bfloat16>());
            k16.back()->alloc(total_pages * page_elems); v16.back()->alloc(total_pages * page_elems);
        } else {
            k32.emplace_back(new DBuf<float>()); v32.emplace_back(new DBuf<float>());
            k32.back()->alloc(total_pages * page_elems); v32.back()->alloc(total_pages * page_elems);
        }
    }
    for (int p = total_pages - 1; p >= 0; --p) free_pages.push_back(p);
    for (int s = B - 1; s >= 0; --s) free_slots.push_back(s);

    // vocoder workspace for the longest chunk
    voc_max_T = CAP;
    {
        const int T1 = (int)std::floor((double)voc_max_T * ((double)c.code_stride / (double)c.output_hop_length));
        const int Tz = (int)std::floor((double)T1 * ((double)c.output_sample_rate / (double)c.input_sample_rate));
        VB = std::max(1, std::min(8, B));       // chunks vocoded per launch (fills the SMs at the short early stages)
        vz.alloc((size_t)VB * c.voc_in_dim * Tz);
        vpre.alloc((size_t)VB * c.voc_init_ch * Tz);
        size_t mx = 0; int len = Tz;
        for (int i = 0; i < c.voc_n_up; ++i) { len *= c.voc_up_rates[i]; mx = std::max(mx, (size_t)stage_ch[i] * len); }
        for (auto& b : vb) b.alloc((size_t)VB * mx);
        vwav.alloc((size_t)VB * len);
        vlat.alloc((size_t)VB * voc_max_T * c.voc_in_dim);
        vcb.alloc((size_t)VB * cbias_stride);
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    launch_base = g_launch_count;
    worker = std::thread([this] { this->loop(); });
}

Engine::~Engine() {
    {
        std::lock_guard<std::mutex> lk(q_mu);
        stop = true;
    }
    cv_work.notify_all();
    if (worker.joinable()) worker.join();
    cudaSetDevice(cfg.device);
    cudaStreamSynchronize(st);
    drop_graphs();
    for (auto& pr : pinned_pool) cudaFreeHost(pr.first);
    for (auto& pr : dev_pool) cudaFree(pr.first);
    for (auto& kv : done_map) {                     // finished chunks nobody fetched
        if (kv.second->wav_host) cudaFreeHost(kv.second->wav_host);
        if (kv.second->wav_dev) cudaFree(kv.second->wav_dev);
        if (kv.second->lat_dev) cudaFree(kv.second->lat_dev);
    }
    if (h_finished) cudaFreeHost(h_finished);
    for (int i = 1; i < kMaxMicro; ++i) { if (st_mb[i]) cudaStreamDestroy(st_mb[i]); if (ev_join[i]) cudaEventDestroy(ev_join[i]); }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_t0) cudaEventDestroy(ev_t0);
    if (ev_t1) cudaEventDestroy(ev_t1);
    if (st) cudaStreamDestroy(st);
}

// ================================================================================================
// weights
// ================================================================================================
void Engine::load_weight(const char* name, const float* data, const int64_t* shape, int ndim) {
    std::lock_guard<std::mutex> lk(mu);
    if (finalized) throw std::runtime_error("weights already finalized");
    HostTensor t;
    t.shape.assign(shape, shape + ndim);
    const size_t n = t.numel();
    t.data.assign(data, data + n);
    raw[name] = std::move(t);
}

const HostTensor& Engine::need(const std::string& name) const {
    auto it = raw.find(name);
    if (it == raw.end()) throw std::runtime_error("missing weight: " + name);
    return it->second;
}

// [K,N] (HF Conv1D, in x out) or [N,K] (nn.Linear) -> device [Npad,K] in the engine precision
void Engine::make_linear(Linear& lin, const std::string& wname, const std::string& bname, bool conv1d_layout, int pad_n) {
    const HostTensor& w = need(wname);
    if (w.shape.size() != 2) throw std::runtime_error("expected 2-D weight: " + wname);
    const int d0 = (int)w.shape[0], d1 = (int)w.shape[1];
    const int N = conv1d_layout ? d1 : d0, K = conv1d_layout ? d0 : d1;
    const int Np = pad_n > 0 ? pad_n : N;
    std::vector<float> t((size_t)Np * K, 0.f);
    if (conv1d_layout) {
        const int BL = 32;                                  // blocked transpose (vllm_mm_gpt.py:723-725)
        for (int k0 = 0; k0 < K; k0 += BL)
            for (int n0 = 0; n0 < N; n0 += BL)
                for (int k = k0; k < std::min(K, k0 + BL); ++k)
                    for (int n = n0; n < std::min(N, n0 + BL); ++n) t[(size_t)n * K + k] = w.data[(size_t)k * N + n];
    } else {
        std::memcpy(t.data(), w.data.data(), (size_t)N * K * sizeof(float));
    }
    lin.N = Np; lin.K = K;
    std::vector<float> bias(Np, 0.f);
    if (!bname.empty()) { const HostTensor& b = need(bname); std::copy(b.data.begin(), b.data.end(), bias.begin()); }
    lin.b.alloc(Np); lin.b.upload(bias.data(), Np, st);
    lin.w32.alloc(t.size()); lin.w32.upload(t.data(), t.size(), st);
    if (bf16) {
        lin.w16.alloc(t.size());
        launch_f32_to_bf16(lin.w32.p, lin.w16.p, t.size(), st);
        CUDA_CHECK(cudaStreamSynchronize(st));
        lin.w32.release();
        weight_bytes += t.size() * 2;
    } else {
        CUDA_CHECK(cudaStreamSynchronize(st));
        weight_bytes += t.size() * 4;
    }
}

// weight-norm fold (torch parametrizations.weight.original0/1, dim 0) — hifigan_decoder.py:44-73,189-202
std::vector<float> Engine::folded(const std::string& prefix) const {
    auto it = raw.find(prefix + ".weight");
    if (it != raw.end()) return it->second.data;
    const HostTensor& g = need(prefix + ".parametrizations.weight.original0");
    const HostTensor& v = need(prefix + ".parametrizations.weight.original1");
    const size_t d0 = (size_t)v.shape[0];
    const size_t inner = v.numel() / d0;
    std::vector<float> w(v.numel());
    for (size_t i = 0; i < d0; ++i) {
        double ss = 0;
        for (size_t j = 0; j < inner; ++j) { const double x = v.data[i * inner + j]; ss += x * x; }
        const float scale = g.data[i] / (float)std::sqrt(ss);
        for (size_t j = 0; j < inner; ++j) w[i * inner + j] = v.data[i * inner + j] * scale;
    }
    return w;
}

// Conv1d weight [Cout,Cin,K] or ConvTranspose1d weight [Cin,Cout,K] -> [Cin][K][Cout]
void Engine::make_conv(ConvW& c, const std::string& prefix, bool transposed, bool has_bias) {
    const std::vector<float> w = folded(prefix);
    const HostTensor& ref = raw.count(prefix + ".weight") ? need(prefix + ".weight")
                                                          : need(prefix + ".parametrizations.weight.original1");
    const int d0 = (int)ref.shape[0], d1 = (int)ref.shape[1], K = (int)ref.shape[2];
    const int Cin = transposed ? d0 : d1, Cout = transposed ? d1 : d0;
    std::vector<float> t((size_t)Cin * K * Cout);
    for (int ci = 0; ci < Cin; ++ci)
        for (int co = 0; co < Cout; ++co)
            for (int j = 0; j < K; ++j) {
                const float x = transposed ? w[((size_t)ci * Cout + co) * K + j] : w[((size_t)co * Cin + ci) * K + j];
                t[((size_t)ci * K + j) * Cout + co] = x;
            }
    c.Cin = Cin; c.Cout = Cout; c.K = K;
    up(c.wt, t);
    if (has_bias) up(c.b, need(prefix + ".bias").data);
    if (bf16 && transposed && K % 2 == 0 && Cout % 32 == 0) {      // fast mode: ConvTranspose1d as u two-tap phases
        const int u = K / 2;
        c.plan = conv1d_tc_plan(Cin, u * Cout, 2);
        if (c.plan.ok) {
            std::vector<__half> blob(c.plan.blob_halves);
            convT_tc_pack(w.data(), Cin, Cout, u, c.plan, blob.data());
            c.blob.alloc(blob.size());
            c.blob.upload(blob.data(), blob.size(), st);
            CUDA_CHECK(cudaStreamSynchronize(st));
            c.tc = true;
            weight_bytes += blob.size() * 2;
        }
    }
    if (bf16 && !transposed) {               // fast mode: fp16 tensor-core tiles for every Conv1d that fits the plan
        c.plan = conv1d_tc_plan(Cin, Cout, K);
        if (c.plan.ok) {
            std::vector<__half> blob(c.plan.blob_halves);
            conv1d_tc_pack(w.data(), Cin, Cout, K, c.plan, blob.data());
            c.blob.alloc(blob.size());
            c.blob.upload(blob.data(), blob.size(), st);
            CUDA_CHECK(cudaStreamSynchronize(st));
            c.tc = true;
            weight_bytes += blob.size() * 2;
        }
    }
}

// Conv1d through whichever path the weights were prepared for
void Engine::conv1d(const ConvW& c, const float* x, const float* cbias, const float* resid, float* out, int L, int dil,
                    float in_scale, float slope, int mode, int nb) {
    launch_conv1d(x, c.wt.p, c.b.p, cbias, resid, out, c.Cin, c.Cout, L, c.K, dil, in_scale, slope, mode, nb, cbias_stride, st);
}

void Engine::conv1d_tc(const ConvW& c, const __half* a16, const float* cbias, const float* resid, float* out32, __half* out16,
                       int L, int lpad, int dil, int mode, int nb, float scale16) {
    launch_conv1d_tc(a16, c.blob.p, c.plan, c.b.p, cbias, resid, out32, out16, c.Cin, c.Cout, L, lpad, c.K, dil, 0.1f, scale16,
                     mode, nb, cbias_stride, st);
}

void Engine::finalize_weights() {
    std::lock_guard<std::mutex> lk(mu);
    if (finalized) return;
    CUDA_CHECK(cudaSetDevice(cfg.device));
    const auto& c = cfg;
    // ---- GPT (names: checkpoint_converter.py:230-272)
    up(text_emb, need("text_embedding.weight").data);
    up(text_pos, need("text_pos_embedding.emb.weight").data);
    up(wte, need("gpt.wte.weight").data);
    up(wpe, need("gpt.wpe.emb.weight").data);
    if ((int)need("gpt.wpe.emb.weight").shape[0] < c.max_audio_tokens + 1) throw std::runtime_error("wpe table too short");
    for (int i = 0; i < L; ++i) {
        const std::string p = "gpt.h." + std::to_string(i) + ".";
        std::unique_ptr<Layer> ly(new Layer());
        up(ly->ln1w, need(p + "ln_1.weight").data); up(ly->ln1b, need(p + "ln_1.bias").data);
        up(ly->ln2w, need(p + "ln_2.weight").data); up(ly->ln2b, need(p + "ln_2.bias").data);
        make_linear(ly->qkv, p + "attn.c_attn.weight", p + "attn.c_attn.bias", true);
        make_linear(ly->o, p + "attn.c_proj.weight", p + "attn.c_proj.bias", true);
        make_linear(ly->fc, p + "mlp.c_fc.weight", p + "mlp.c_fc.bias", true);
        make_linear(ly->proj, p + "mlp.c_proj.weight", p + "mlp.c_proj.bias", true);
        if (ly->qkv.N != 3 * H || ly->qkv.K != H || ly->fc.N != FF || ly->proj.K != FF) throw std::runtime_error("GPT weight shape mismatch at layer " + std::to_string(i));
        layers.push_back(std::move(ly));
        // free host copies early (1.5 GB for the full model)
        for (const char* nm : {"attn.c_attn.weight", "attn.c_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight"}) raw.erase(p + nm);
    }
    up(lnfw, need("gpt.ln_f.weight").data); up(lnfb, need("gpt.ln_f.bias").data);
    up(fnw, need("final_norm.weight").data); up(fnb, need("final_norm.bias").data);
    make_linear(mel_head, "mel_head.weight", "mel_head.bias", false, Vpad);
    // ---- vocoder (names: XTTSv2Engine.state_dict(), hifigan_decoder.*)
    const std::string w = "hifigan_decoder.waveform_decoder.";
    make_conv(conv_pre, w + "conv_pre", false, true);
    {
        const HostTensor& cw = need(w + "cond_layer.weight");
        cond_layer.rows = (int)cw.shape[0];
        up(cond_layer.w, cw.data); up(cond_layer.b, need(w + "cond_layer.bias").data);
    }
    const int nk = c.voc_n_rb;
    for (int i = 0; i < c.voc_n_up; ++i) {
        std::unique_ptr<ConvW> u(new ConvW());
        make_conv(*u, w + "ups." + std::to_string(i), true, true);
        ups.push_back(std::move(u));
        std::unique_ptr<CondLin> cl(new CondLin());
        const HostTensor& cw = need(w + "conds." + std::to_string(i) + ".weight");
        cl->rows = (int)cw.shape[0];
        up(cl->w, cw.data); up(cl->b, need(w + "conds." + std::to_string(i) + ".bias").data);
        conds.push_back(std::move(cl));
        for (int j = 0; j < nk; ++j) {
            std::unique_ptr<RB> rb(new RB());
            const std::string rp = w + "resblocks." + std::to_string(i * nk + j) + ".";
            for (int t = 0; t < 3; ++t) {
                rb->c1[t].reset(new ConvW()); rb->c2[t].reset(new ConvW());
                make_conv(*rb->c1[t], rp + "convs1." + std::to_string(t), false, true);
                make_conv(*rb->c2[t], rp + "convs2." + std::to_string(t), false, true);
            }
            rbs.push_back(std::move(rb));
        }
    }
    {
        const HostTensor& pw = need(w + "conv_post.weight");
        post_cin = (int)pw.shape[1];
        up(conv_post_w, pw.data);
    }
    if (bf16) {
        bool all_tc = conv_pre.tc;
        for (auto& u : ups) all_tc = all_tc && u->tc && (u->K == 2 * c.voc_up_rates[&u - &ups[0]]);
        for (auto& rb : rbs) for (int t = 0; t < 3; ++t) all_tc = all_tc && rb->c1[t]->tc && rb->c2[t]->tc;
        for (int ch : stage_ch) all_tc = all_tc && (ch % 16 == 0);
        if (all_tc) {
            const int T1 = (int)std::floor((double)voc_max_T * ((double)c.code_stride / (double)c.output_hop_length));
            const int Tz = (int)std::floor((double)T1 * ((double)c.output_sample_rate / (double)c.input_sample_rate));
            vz16.alloc((size_t)VB * c.voc_in_dim * atoms_lpad(Tz));
            size_t mx = (size_t)c.voc_init_ch * atoms_lpad(Tz); int len = Tz;
            for (int i = 0; i < c.voc_n_up; ++i) { len *= c.voc_up_rates[i]; mx = std::max(mx, (size_t)stage_ch[i] * atoms_lpad(len)); }
            for (auto& b : va16) b.alloc((size_t)VB * mx);
            tc_vocoder_ready = true;
        }
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    // ---- speaker conditioning stack (optional in a checkpoint: without it only xtts_set_speaker works)
    if (raw.count("conditioning_encoder.init.weight") && raw.count("hifigan_decoder.speaker_encoder.conv1.weight")) {
        auto get = [this](const std::string& n) {
            const HostTensor& t = need(n);
            return HostTensorView{t.data.data(), t.shape};
        };
        conditioner.reset(new Conditioner(cfg, get, st));
    }
    raw.clear();
    finalized = true;
}

// ================================================================================================
// speakers
// ================================================================================================
void Engine::set_speaker(int slot, const float* cond, const float* g) {
    std::lock_guard<std::mutex> lk(mu);
    require_finalized();
    if (slot < 0 || slot >= S) throw std::runtime_error("speaker slot out of range");
    CUDA_CHECK(cudaSetDevice(cfg.device));
    const size_t nc = (size_t)cfg.n_cond_latents * H;
    spk_cond.upload(cond, nc, st, (size_t)slot * nc);
    spk_g.upload(g, cfg.d_vector, st, (size_t)slot * cfg.d_vector);
    // speaker-conditioning biases: cond_layer(g), conds[i](g)  (hifigan_decoder.py:244-251)
    finish_speaker(slot);
}

void Engine::finish_speaker(int slot) {
    const float* gd = spk_g.p + (size_t)slot * cfg.d_vector;
    float* cb = spk_cbias.p + (size_t)slot * cbias_stride;
    launch_gemv(cond_layer.w.p, cond_layer.b.p, gd, cb + cbias_off[0], cond_layer.rows, cfg.d_vector, st);
    for (int i = 0; i < cfg.voc_n_up; ++i)
        launch_gemv(conds[i]->w.p, conds[i]->b.p, gd, cb + cbias_off[i + 1], conds[i]->rows, cfg.d_vector, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    spk_valid[slot] = 1;
}

// get_conditioning_latents (XTTSv2.py:409-468) on the GPU
void Engine::condition(int slot, const float* w22, int64_t n22, const float* w16, int64_t n16, int cond_len, int chunk_len) {
    std::lock_guard<std::mutex> lk(mu);
    require_finalized();
    if (!conditioner) throw std::runtime_error("checkpoint has no conditioning encoder / speaker encoder weights");
    if (slot < 0 || slot >= S) throw std::runtime_error("speaker slot out of range");
    if (cfg.spk_proj != cfg.d_vector) throw std::runtime_error("speaker encoder projection != d_vector");
    CUDA_CHECK(cudaSetDevice(cfg.device));
    const double t0 = now_s();
    conditioner->run(w22, n22, w16, n16, cond_len, chunk_len, spk_cond.p + (size_t)slot * cfg.n_cond_latents * H,
                     spk_g.p + (size_t)slot * cfg.d_vector);
    finish_speaker(slot);
    st_cond_ms += (now_s() - t0) * 1e3;
}

void Engine::get_speaker(int slot, float* cond, float* g) {
    std::lock_guard<std::mutex> lk(mu);
    if (slot < 0 || slot >= S || !spk_valid[slot]) throw std::runtime_error("speaker slot not set");
    CUDA_CHECK(cudaSetDevice(cfg.device));
    const size_t nc = (size_t)cfg.n_cond_latents * H;
    if (cond) spk_cond.download(cond, nc, st, (size_t)slot * nc);
    if (g) spk_g.download(g, cfg.d_vector, st, (size_t)slot * cfg.d_vector);
    CUDA_CHECK(cudaStreamSynchronize(st));
}

// ================================================================================================
// GPT forward
// ================================================================================================
SampleState Engine::sample_state() const {
    SampleState s;
    s.last_tok = d_last_tok.p; s.n_gen = d_n_gen.p; s.ctx_len = d_ctx_len.p; s.finished = d_finished.p;
    s.tokens = d_tokens.p; s.sampled = d_sampled.p; s.forced = use_forced ? d_forced.p : nullptr;
    s.seen = d_seen.p; s.temperature = d_temp.p; s.top_p = d_top_p.p; s.top_k = d_top_k.p; s.penalty = d_pen.p;
    s.max_tokens = d_max_tokens.p; s.stop_token = d_stop.p; s.seed = d_seed.p; s.seq_seed = d_seq_seed.p;
    s.tokens_cap = CAP; s.seen_words = SEENW;
    return s;
}

void Engine::gemm(const void* A, const Linear& lin, const float* resid, void* out, int M, int flags, bool pdl) {
    if (bf16)
        launch_gemm_bf16_tc(reinterpret_cast<const __nv_bfloat16*>(A), lin.w16.p, lin.b.p, resid, out, M, lin.N, lin.K, flags, st, pdl);
    else
        launch_gemm_f32(reinterpret_cast<const float*>(A), lin.w32.p, lin.b.p, resid, reinterpret_cast<float*>(out), M, lin.N, lin.K,
                        flags & ~GEMM_OUT_BF16, st);
}

// X [M,H] -> X after all blocks.  prefill: causal attention inside each sequence of d_attnseq, KV written at
// d_row_pos; decode: one row per active slot, KV appended at ctx_len, attention over the paged cache.
void Engine::layers_forward(int M, bool prefill, int nseq, int max_nq) {
    void* Xn = bf16 ? (void*)wXn16.p : (void*)wXn32.p;
    void* ATT = bf16 ? (void*)wATT16.p : (void*)wATT32.p;
    void* FFb = bf16 ? (void*)wFF16.p : (void*)wFF32.p;
    const int oflag = bf16 ? GEMM_OUT_BF16 : 0;
    // Decode-shaped steps in fast mode: the two N = hidden GEMMs (attention out-proj, MLP down-proj) expose only
    // N/BN CTAs, so they run split-K into fp32 partials and the reduction is fused with the residual add and the
    // following LayerNorm (fixed summation order => deterministic).
    const bool splitk = bf16 && !prefill && use_splitk && M <= NSLOT && (H / 64) % 4 == 0 && (FF / 64) % 8 == 0;
    const bool pdl = !prefill && use_pdl;            // decode chain: programmatic dependent launch
    auto ln = [&](const float* w, const float* b) {
        if (bf16) launch_layernorm<__nv_bfloat16>(wX.p, w, b, wXn16.p, M, H, cfg.ln_eps, st, pdl);
        else launch_layernorm<float>(wX.p, w, b, wXn32.p, M, H, cfg.ln_eps, st, pdl);
    };
    if (splitk) ln(layers[0]->ln1w.p, layers[0]->ln1b.p);
    for (int l = 0; l < L; ++l) {
        Layer& ly = *layers[l];
        if (!splitk) ln(ly.ln1w.p, ly.ln1b.p);
        gemm(Xn, ly.qkv, nullptr, wQKV.p, M, 0, pdl);
        if (prefill) {
            if (bf16) launch_kv_write<__nv_bfloat16>(wQKV.p, M, d_row_slot.p, d_row_pos.p, d_ctx_len.p, d_block_tables.p, max_pages, k16[l]->p, v16[l]->p, NH, st);
            else launch_kv_write<float>(wQKV.p, M, d_row_slot.p, d_row_pos.p, d_ctx_len.p, d_block_tables.p, max_pages, k32[l]->p, v32[l]->p, NH, st);
            AttnLayout A;
            A.q = wQKV.p; A.k = wQKV.p + H; A.v = wQKV.p + 2 * H;
            A.q_row_stride = 3 * H; A.kv_row_stride = 3 * H; A.q_head_stride = kHeadDim; A.kv_head_stride = kHeadDim;
            A.heads = NH; A.scale = 0.125f; A.causal = 1;
            if (bf16) launch_attn_generic<__nv_bfloat16>(A, d_attnseq.p, nseq, max_nq, wATT16.p, H, st);
            else launch_attn_generic<float>(A, d_attnseq.p, nseq, max_nq, wATT32.p, H, st);
        } else {
            // (the attention kernel appends this step's K/V to the cache itself)
            if (bf16) launch_attn_decode<__nv_bfloat16, __nv_bfloat16>(wQKV.p, d_active.p, M, d_ctx_len.p, d_block_tables.p, max_pages, k16[l]->p, v16[l]->p, wATT16.p, NH, st, decode_ctx_sum, pdl);
            else launch_attn_decode<float, float>(wQKV.p, d_active.p, M, d_ctx_len.p, d_block_tables.p, max_pages, k32[l]->p, v32[l]->p, wATT32.p, NH, st, decode_ctx_sum, pdl);
        }
        if (splitk) {
            launch_gemm_bf16_tc_splitk(wATT16.p, ly.o.w16.p, wPART.p, M, H, H, 4, st, pdl);
            launch_residual_reduce_layernorm<__nv_bfloat16>(wX.p, wPART.p, 4, ly.o.b.p, ly.ln2w.p, ly.ln2b.p, wXn16.p, M, H, cfg.ln_eps, st, pdl);
            gemm(Xn, ly.fc, nullptr, FFb, M, GEMM_GELU | oflag, pdl);
            launch_gemm_bf16_tc_splitk(wFF16.p, ly.proj.w16.p, wPART.p, M, H, FF, 8, st, pdl);
            const bool last = (l + 1 == L);
            launch_residual_reduce_layernorm<__nv_bfloat16>(wX.p, wPART.p, 8, ly.proj.b.p, last ? nullptr : layers[l + 1]->ln1w.p,
                                                            last ? nullptr : layers[l + 1]->ln1b.p, last ? nullptr : wXn16.p, M, H,
                                                            cfg.ln_eps, st, pdl);
        } else {
            gemm(ATT, ly.o, wX.p, wX.p, M, GEMM_RESID, pdl);
            ln(ly.ln2w.p, ly.ln2b.p);
            gemm(Xn, ly.fc, nullptr, FFb, M, GEMM_GELU | oflag, pdl);
            gemm(FFb, ly.proj, wX.p, wX.p, M, GEMM_RESID, pdl);
        }
    }
}

// rows row_index[0..M) of X -> Y -> logits (wLOG[i]) ; latents captured ; optionally sample
void Engine::head_and_sample(int M, const int* row_index, const int* slots_dev, const int* lat_pos, int advance_ctx, bool do_sample,
                             bool pdl_first) {
    const bool pdl = advance_ctx && use_pdl;         // decode step only
    const bool pdl0 = pdl && pdl_first;              // (the first kernel after a stream join takes a full dependency)
    if (bf16) launch_head_norms<__nv_bfloat16>(wX.p, row_index, lnfw.p, lnfb.p, fnw.p, fnb.p, wY16.p, d_latents.p, slots_dev, lat_pos, d_n_gen.p, CAP, M, H, cfg.ln_eps, st, pdl0);
    else launch_head_norms<float>(wX.p, row_index, lnfw.p, lnfb.p, fnw.p, fnb.p, wY32.p, d_latents.p, slots_dev, lat_pos, d_n_gen.p, CAP, M, H, cfg.ln_eps, st, pdl0);
    gemm(bf16 ? (void*)wY16.p : (void*)wY32.p, mel_head, nullptr, wLOG.p, M, 0, pdl);
    if (do_sample) launch_sample(wLOG.p, Vpad, slots_dev, M, V, sample_state(), advance_ctx, st, pdl);
}

void Engine::init_slot(Sequence& s, const int32_t* forced, int n_forced) {
    const int slot = s.slot;
    const int n_text = (int)s.text_ids.size();
    s.n_prompt = cfg.n_cond_latents + n_text + 1;
    const int max_tok = std::min<int>(s.sp.max_tokens > 0 ? s.sp.max_tokens : CAP, CAP);
    const int need_pages = ceil_div(s.n_prompt + max_tok, kPageTokens);
    if ((int)free_pages.size() < need_pages) throw std::runtime_error("out of KV pages");
    s.pages.clear();
    std::vector<int> bt(max_pages, 0);
    for (int i = 0; i < need_pages; ++i) { bt[i] = free_pages.back(); free_pages.pop_back(); s.pages.push_back(bt[i]); }
    d_block_tables.upload(bt.data(), max_pages, st, (size_t)slot * max_pages);
    const int zero = 0, ctx = s.n_prompt;
    d_n_gen.upload(&zero, 1, st, slot); d_finished.upload(&zero, 1, st, slot); d_ctx_len.upload(&ctx, 1, st, slot);
    d_last_tok.upload(&cfg.start_audio_token, 1, st, slot);
    const int tk = s.sp.top_k, stopt = s.sp.stop_token, ss = s.sp.seq_seed;
    const float T = s.sp.temperature, tp = s.sp.top_p, pen = s.sp.repetition_penalty;
    const unsigned long long seed = s.sp.seed;
    d_top_k.upload(&tk, 1, st, slot); d_max_tokens.upload(&max_tok, 1, st, slot); d_stop.upload(&stopt, 1, st, slot);
    d_seq_seed.upload(&ss, 1, st, slot); d_temp.upload(&T, 1, st, slot); d_top_p.upload(&tp, 1, st, slot);
    d_pen.upload(&pen, 1, st, slot); d_seed.upload(&seed, 1, st, slot);
    // penalty set seed: prompt ids are [1]*(32+Lt)+[start]  (vllm_mm_gpt.py:325, App. B.7)
    std::vector<unsigned> seen(SEENW, 0u);
    seen[1 >> 5] |= 1u << 1;
    seen[cfg.start_audio_token >> 5] |= 1u << (cfg.start_audio_token & 31);
    d_seen.upload(seen.data(), SEENW, st, (size_t)slot * SEENW);
    std::vector<int> f(CAP, -1);
    if (forced) for (int i = 0; i < std::min(n_forced, CAP); ++i) f[i] = forced[i];
    d_forced.upload(f.data(), CAP, st, (size_t)slot * CAP);
    CUDA_CHECK(cudaStreamSynchronize(st));     // host staging vectors go out of scope
}

void Engine::release_slot(Sequence& s) {
    for (int p : s.pages) free_pages.push_back(p);
    s.pages.clear();
    if (s.slot >= 0 && s.slot < B) free_slots.push_back(s.slot);
    s.slot = -1;
}

// builds row descriptors for [prompt ; optional forced audio rows] of each sequence; returns total rows
int Engine::build_prefill(const std::vector<Sequence*>& seqs, const std::vector<std::vector<int32_t>>& audio,
                          std::vector<int>& last_rows, int& max_nq) {
    std::vector<RowDesc> rows;
    std::vector<int> row_slot, row_pos;
    std::vector<AttnSeq> as;
    last_rows.clear();
    max_nq = 0;
    for (size_t si = 0; si < seqs.size(); ++si) {
        Sequence& s = *seqs[si];
        const int start = (int)rows.size();
        for (int i = 0; i < cfg.n_cond_latents; ++i) rows.push_back(RowDesc{0, i, 0, s.speaker});
        for (int i = 0; i < (int)s.text_ids.size(); ++i) {
            const int id = s.text_ids[i];
            if (id < 0 || id >= cfg.n_text_tokens) throw std::runtime_error("text token id out of range");
            rows.push_back(RowDesc{1, id, i, 0});
        }
        rows.push_back(RowDesc{2, cfg.start_audio_token, 0, 0});
        if (si < audio.size())
            for (int k = 0; k < (int)audio[si].size(); ++k) {
                const int id = audio[si][k];
                if (id < 0 || id >= V) throw std::runtime_error("audio token id out of range");
                rows.push_back(RowDesc{2, id, k + 1, 0});
            }
        const int n = (int)rows.size() - start;
        for (int i = 0; i < n; ++i) { row_slot.push_back(s.slot); row_pos.push_back(i); }
        as.push_back(AttnSeq{start, n, start, n});
        last_rows.push_back(start + n - 1);
        max_nq = std::max(max_nq, n);
    }
    const int M = (int)rows.size();
    if (M > prefill_rows_cap) throw std::runtime_error("prefill batch exceeds row budget");
    d_rows.upload(rows.data(), M, st); d_row_slot.upload(row_slot.data(), M, st); d_row_pos.upload(row_pos.data(), M, st);
    d_attnseq.upload(as.data(), as.size(), st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    return M;
}

void Engine::prefill(const std::vector<Sequence*>& seqs) {
    std::vector<int> last_rows; int max_nq = 0;
    const int M = build_prefill(seqs, {}, last_rows, max_nq);
    launch_build_rows(d_rows.p, M, tables(), wX.p, st);
    layers_forward(M, true, (int)seqs.size(), max_nq);
    std::vector<int> slots;
    for (auto* s : seqs) slots.push_back(s->slot);
    d_rowidx.upload(last_rows.data(), last_rows.size(), st);
    d_active.upload(slots.data(), slots.size(), st);
    head_and_sample((int)seqs.size(), d_rowidx.p, d_active.p, nullptr, 0, true);
    CUDA_CHECK(cudaStreamSynchronize(st));
    st_prefill_rows += M;
    const double t = now_s();
    for (auto* s : seqs) s->t_first = t;
}

// Fast-mode decode layers for rows [r0, r0 + Mi) of the step on stream `s`.  Every work buffer is row-major and the
// KV cache is per slot, so disjoint row ranges are independent: decode_step runs several of these as concurrent
// branches (micro-batches), which lets one branch's HBM-bound attention overlap another's latency-bound GEMM chain.
void Engine::decode_layers_rows(int r0, int Mi, cudaStream_t s, bool pdl_first, double ctx_sum) {
    float* X = wX.p + (size_t)r0 * H;
    __nv_bfloat16* Xn = wXn16.p + (size_t)r0 * H;
    float* QKV = wQKV.p + (size_t)r0 * 3 * H;
    __nv_bfloat16* ATT = wATT16.p + (size_t)r0 * H;
    __nv_bfloat16* FFb = wFF16.p + (size_t)r0 * FF;
    float* PART = wPART.p + (size_t)r0 * 8 * H;          // [splits <= 8][Mi][H] inside this branch's own region
    const int* act = d_active.p + r0;
    const bool pdl = use_pdl;
    launch_layernorm<__nv_bfloat16>(X, layers[0]->ln1w.p, layers[0]->ln1b.p, Xn, Mi, H, cfg.ln_eps, s, pdl && pdl_first);
    for (int l = 0; l < L; ++l) {
        Layer& ly = *layers[l];
        launch_gemm_bf16_tc(Xn, ly.qkv.w16.p, ly.qkv.b.p, nullptr, QKV, Mi, ly.qkv.N, ly.qkv.K, 0, s, pdl);
        launch_attn_decode<__nv_bfloat16, __nv_bfloat16>(QKV, act, Mi, d_ctx_len.p, d_block_tables.p, max_pages, k16[l]->p, v16[l]->p,
                                                         ATT, NH, s, ctx_sum, pdl);
        launch_gemm_bf16_tc_splitk(ATT, ly.o.w16.p, PART, Mi, H, H, 4, s, pdl);
        launch_residual_reduce_layernorm<__nv_bfloat16>(X, PART, 4, ly.o.b.p, ly.ln2w.p, ly.ln2b.p, Xn, Mi, H, cfg.ln_eps, s, pdl);
        launch_gemm_bf16_tc(Xn, ly.fc.w16.p, ly.fc.b.p, nullptr, FFb, Mi, ly.fc.N, ly.fc.K, GEMM_GELU | GEMM_OUT_BF16, s, pdl);
        launch_gemm_bf16_tc_splitk(FFb, ly.proj.w16.p, PART, Mi, H, FF, 8, s, pdl);
        const bool last = (l + 1 == L);
        launch_residual_reduce_layernorm<__nv_bfloat16>(X, PART, 8, ly.proj.b.p, last ? nullptr : layers[l + 1]->ln1w.p,
                                                        last ? nullptr : layers[l + 1]->ln1b.p, last ? nullptr : Xn, Mi, H,
                                                        cfg.ln_eps, s, pdl);
    }
}

// Fast-mode decode layers with the fused chain kernel: per layer one attention launch and one persistent launch that
// runs out-proj, residual+LN2, fc+gelu, down-proj, residual+LN1(next) and the next layer's QKV projection.
void Engine::decode_layers_chain(int M) {
    const bool pdl = use_pdl;
    launch_layernorm<__nv_bfloat16>(wX.p, layers[0]->ln1w.p, layers[0]->ln1b.p, wXn16.p, M, H, cfg.ln_eps, st, pdl);
    launch_gemm_bf16_tc(wXn16.p, layers[0]->qkv.w16.p, layers[0]->qkv.b.p, nullptr, wQKV.p, M, layers[0]->qkv.N, layers[0]->qkv.K, 0, st, pdl);
    for (int l = 0; l < L; ++l) {
        Layer& ly = *layers[l];
        Layer* nx = (l + 1 < L) ? layers[l + 1].get() : nullptr;
        launch_attn_decode<__nv_bfloat16, __nv_bfloat16>(wQKV.p, d_active.p, M, d_ctx_len.p, d_block_tables.p, max_pages, k16[l]->p, v16[l]->p,
                                                         wATT16.p, NH, st, decode_ctx_sum, pdl);
        DecodeChainArgs a{};
        a.phases = DC_PROJ | DC_LN2 | DC_FC | DC_FC2 | DC_LN1 | (nx ? DC_QKV : 0);
        a.M = M; a.H = H; a.FF = FF; a.eps = cfg.ln_eps;
        a.ATT = wATT16.p; a.Wo = ly.o.w16.p; a.Wfc = ly.fc.w16.p; a.Wproj = ly.proj.w16.p; a.Wqkv = nx ? nx->qkv.w16.p : nullptr;
        a.proj_bias = ly.o.b.p; a.ln2_w = ly.ln2w.p; a.ln2_b = ly.ln2b.p; a.fc_bias = ly.fc.b.p; a.fc2_bias = ly.proj.b.p;
        a.ln1_w = nx ? nx->ln1w.p : nullptr; a.ln1_b = nx ? nx->ln1b.p : nullptr; a.qkv_bias = nx ? nx->qkv.b.p : nullptr;
        a.X = wX.p; a.Xn = wXn16.p; a.FFb = wFF16.p; a.QKV = wQKV.p; a.PART = wPART.p; a.sync = d_chain_sync.p;
        launch_decode_chain(a, st, pdl);
    }
}

void Engine::drop_graphs() {
    for (auto& kv : decode_graphs) cudaGraphExecDestroy(kv.second);
    decode_graphs.clear();
    for (auto& kv : prof_graphs) {
        cudaGraphExecDestroy(kv.second.first);
        for (auto& r : kv.second.second->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        delete kv.second.second;
    }
    prof_graphs.clear();
    last_prof = nullptr;
}

void Engine::decode_step(const std::vector<int>& active) {
    const int M = (int)active.size();
    d_active.upload(active.data(), M, st);
    const bool fast = bf16 && use_splitk && M <= NSLOT && (H / 64) % 4 == 0 && (FF / 64) % 8 == 0;
    const bool chain = fast && use_chain && decode_chain_supported(M, H, FF);
    const int nmb = (fast && !chain && n_micro > 1 && M >= micro_min_rows) ? std::min(n_micro, (int)kMaxMicro) : 1;
    auto enqueue = [&] {
        launch_build_decode_rows(d_active.p, M, d_last_tok.p, d_n_gen.p, tables(), wX.p, st, use_pdl);
        if (nmb > 1) {
            // fork: every branch starts after the row build; join: head + sampler run once over all rows
            CUDA_CHECK(cudaEventRecord(ev_fork, st));
            for (int i = 1; i < nmb; ++i) CUDA_CHECK(cudaStreamWaitEvent(st_mb[i], ev_fork, 0));
            for (int i = 0, r0 = 0; i < nmb; ++i) {
                const int Mi = M / nmb + (i < M % nmb ? 1 : 0);
                decode_layers_rows(r0, Mi, st_mb[i], i == 0, decode_ctx_sum * (double)Mi / (double)M);
                r0 += Mi;
            }
            for (int i = 1; i < nmb; ++i) {
                CUDA_CHECK(cudaEventRecord(ev_join[i], st_mb[i]));
                CUDA_CHECK(cudaStreamWaitEvent(st, ev_join[i], 0));
            }
            head_and_sample(M, nullptr, d_active.p, nullptr, 1, true, false);
        } else {
            if (chain) decode_layers_chain(M); else layers_forward(M, false, 0, 0);
            head_and_sample(M, nullptr, d_active.p, nullptr, 1, true);
        }
    };
    // The decode step is ~250 small launches whose arguments depend only on M (slot lists, positions and
    // lengths live in device memory), so it is captured once per batch size into a CUDA graph and replayed.
    // Kernel-family profiling and teacher forcing use the eager path.
    const bool graphable = use_graphs && !use_forced && eager_steps_done >= 2;
    if (!graphable) {
        enqueue();
        ++eager_steps_done;
    } else if (g_prof.enabled) {
        // kernel-family profiling: the same step captured with an event-record node on either side of every kernel
        // (full dependencies instead of PDL edges), replayed, and read back after the step's synchronize
        auto it = prof_graphs.find(M);
        if (it == prof_graphs.end()) {
            cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr;
            auto* recs = new KernelProfiler::GraphRecs();
            recs->ctx_sum = decode_ctx_sum;
            const unsigned long long lc = g_launch_count;
            const bool pdl_was = g_use_pdl;
            g_use_pdl = false; g_prof.cap = recs;
            CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try { enqueue(); }
            catch (...) { g_prof.cap = nullptr; g_use_pdl = pdl_was; cudaStreamEndCapture(st, &g); if (g) cudaGraphDestroy(g); delete recs; throw; }
            g_prof.cap = nullptr; g_use_pdl = pdl_was;
            CUDA_CHECK(cudaStreamEndCapture(st, &g));
            CUDA_CHECK(cudaGraphInstantiate(&ge, g, 0));
            cudaGraphDestroy(g);
            graph_kernels[M] = g_launch_count - lc;
            g_launch_count = lc;
            it = prof_graphs.emplace(M, std::make_pair(ge, recs)).first;
        }
        CUDA_CHECK(cudaGraphLaunch(it->second.first, st));
        g_launch_count += graph_kernels[M];
        last_prof = it->second.second;
    } else {
        auto it = decode_graphs.find(M);
        if (it == decode_graphs.end()) {
            cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr;
            const unsigned long long lc = g_launch_count;
            CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try { enqueue(); }
            catch (...) { cudaStreamEndCapture(st, &g); if (g) cudaGraphDestroy(g); throw; }
            CUDA_CHECK(cudaStreamEndCapture(st, &g));
            CUDA_CHECK(cudaGraphInstantiate(&ge, g, 0));
            cudaGraphDestroy(g);
            graph_kernels[M] = g_launch_count - lc;
            g_launch_count = lc;                       // capture enqueued nothing on the device
            it = decode_graphs.emplace(M, ge).first;
        }
        CUDA_CHECK(cudaGraphLaunch(it->second, st));
        g_launch_count += graph_kernels[M];            // kernels executed by the replay
    }
    ++st_decode_steps;
}

// ================================================================================================
// vocoder driver  (HifiDecoder.forward, hifigan_decoder.py:776-802 + HifiganGenerator.forward :228-260)
// ================================================================================================
// `nb` equal-length chunks at once: latents [nb][T][in_dim] (contiguous, device) -> wav_dev_out [nb][n_samples]
void Engine::run_vocoder(const float* lat_dev, int T, const int* speakers, int nb, float* wav_dev_out, int* n_out,
                         const char* stage, float* stage_out, int64_t stage_cap) {
    const auto& c = cfg;
    if (T <= 0 || T > voc_max_T) throw std::runtime_error("vocoder: latent count out of range");
    if (nb < 1 || nb > VB) throw std::runtime_error("vocoder: batch out of range");
    for (int i = 0; i < nb; ++i) {
        const int sp = speakers[i];
        if (sp < 0 || sp >= S || !spk_valid[sp]) throw std::runtime_error("vocoder: speaker slot not set");
        CUDA_CHECK(cudaMemcpyAsync(vcb.p + (size_t)i * cbias_stride, spk_cbias.p + (size_t)sp * cbias_stride,
                                   (size_t)cbias_stride * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (tc_vocoder_ready && use_tc_vocoder) { run_vocoder_tc(lat_dev, T, nb, wav_dev_out, n_out, stage, stage_out, stage_cap); return; }
    const double s1 = (double)c.code_stride / (double)c.output_hop_length;
    const double s2 = (double)c.output_sample_rate / (double)c.input_sample_rate;
    const int T1 = (int)std::floor((double)T * s1);
    const bool resample = c.output_sample_rate != c.input_sample_rate;
    const int Tz = resample ? (int)std::floor((double)T1 * s2) : T1;
    const float* cb = vcb.p;
    auto dump = [&](const char* name, const float* p, size_t n) {           // first batch item only
        if (stage && stage_out && std::strcmp(stage, name) == 0) {
            const size_t m = std::min<size_t>(n, (size_t)stage_cap);
            CUDA_CHECK(cudaMemcpyAsync(stage_out, p, m * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
    };
    launch_interp(lat_dev, vz.p, nullptr, 0, T, c.voc_in_dim, T1, Tz, s1, resample ? s2 : 1.0, nb, st);
    dump("z", vz.p, (size_t)c.voc_in_dim * Tz);
    conv1d(conv_pre, vz.p, cb + cbias_off[0], nullptr, vpre.p, Tz, 1, 1.0f, 1.0f, CONV_STORE, nb);
    dump("pre", vpre.p, (size_t)c.voc_init_ch * Tz);
    const float* cur = vpre.p;
    float in_scale = 1.0f;
    int len = Tz;
    const int nk = c.voc_n_rb;
    float* X = vb[0].p; float* TMP = vb[1].p; float* R1 = vb[2].p; float* R2 = vb[3].p; float* ZS = vb[4].p;
    for (int i = 0; i < c.voc_n_up; ++i) {
        const ConvW& u = *ups[i];
        launch_conv_transpose1d(cur, u.wt.p, u.b.p, cb + cbias_off[i + 1], X, nullptr, 0, 0.f, u.Cin, u.Cout, len, u.K,
                                c.voc_up_rates[i], in_scale, 0.1f, nb, cbias_stride, st);
        len *= c.voc_up_rates[i];
        const int C = u.Cout;
        { char nm[16]; snprintf(nm, sizeof(nm), "up%d", i); dump(nm, X, (size_t)C * len); }
        for (int j = 0; j < nk; ++j) {
            const RB& rb = *rbs[i * nk + j];
            const float* r_in = X;
            for (int t = 0; t < 3; ++t) {
                const ConvW& a = *rb.c1[t];
                const ConvW& b = *rb.c2[t];
                conv1d(a, r_in, nullptr, nullptr, TMP, len, c.voc_rb_dilations[t], 1.0f, 0.1f, CONV_STORE, nb);
                if (t < 2) {
                    float* r_out = (t == 0) ? R1 : R2;
                    conv1d(b, TMP, nullptr, r_in, r_out, len, 1, 1.0f, 0.1f, CONV_STORE, nb);
                    r_in = r_out;
                } else {
                    conv1d(b, TMP, nullptr, r_in, ZS, len, 1, 1.0f, 0.1f, j == 0 ? CONV_STORE : CONV_ACCUM, nb);
                }
            }
        }
        { char nm[16]; snprintf(nm, sizeof(nm), "mrf%d", i); dump(nm, ZS, (size_t)C * len); }   // un-normalised sum
        // next stage reads the MRF sum scaled by 1/nk; its ConvT writes X (dead by now), and ZS is only
        // overwritten after that ConvT has consumed it (stream order)
        cur = ZS;
        in_scale = 1.0f / (float)nk;
    }
    launch_conv_post(cur, conv_post_w.p, wav_dev_out, post_cin, len, 7, in_scale, 0.01f, nb, st);
    *n_out = len;
}

// Tensor-core vocoder: every Conv1d operand is kept as activated fp16 atoms written by its producer's epilogue
// (ConvT / previous conv), so the conv kernels are pure bulk-copy + tcgen05; fp32 is kept for the residual stream
// (x, r1, r2), the MRF sum and the final waveform.  (speaker biases were gathered into vcb by run_vocoder)
void Engine::run_vocoder_tc(const float* lat_dev, int T, int nb, float* wav_dev_out, int* n_out, const char* stage,
                            float* stage_out, int64_t stage_cap) {
    const auto& c = cfg;
    const double s1 = (double)c.code_stride / (double)c.output_hop_length;
    const double s2 = (double)c.output_sample_rate / (double)c.input_sample_rate;
    const int T1 = (int)std::floor((double)T * s1);
    const bool resample = c.output_sample_rate != c.input_sample_rate;
    const int Tz = resample ? (int)std::floor((double)T1 * s2) : T1;
    const float* cb = vcb.p;
    auto dump = [&](const char* name, const float* p, size_t n) {
        if (stage && stage_out && std::strcmp(stage, name) == 0) {
            const size_t m = std::min<size_t>(n, (size_t)stage_cap);
            CUDA_CHECK(cudaMemcpyAsync(stage_out, p, m * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
    };
    int lpad = atoms_lpad(Tz);
    __half* XA = va16[0].p; __half* TA = va16[1].p; __half* RA[2] = {va16[2].p, va16[3].p}; __half* PA = va16[4].p;
    launch_atoms_zero_pads(vz16.p, nb * c.voc_in_dim / 8, lpad, Tz, st);
    launch_atoms_zero_pads(PA, nb * c.voc_init_ch / 8, lpad, Tz, st);
    launch_interp(lat_dev, stage ? vz.p : nullptr, vz16.p, lpad, T, c.voc_in_dim, T1, Tz, s1, resample ? s2 : 1.0, nb, st);
    if (stage) dump("z", vz.p, (size_t)c.voc_in_dim * Tz);
    // conv_pre: fp32 copy only for the stage tap; its activated fp16 atoms feed the first transposed conv
    conv1d_tc(conv_pre, vz16.p, cb + cbias_off[0], nullptr, stage ? vpre.p : nullptr, PA, Tz, lpad, 1, CONV_STORE, nb);
    if (stage) dump("pre", vpre.p, (size_t)c.voc_init_ch * Tz);
    int len = Tz;
    const int nk = c.voc_n_rb;
    float* X = vb[0].p; float* R[2] = {vb[2].p, vb[3].p}; float* ZS = vb[4].p;
    for (int i = 0; i < c.voc_n_up; ++i) {
        const ConvW& u = *ups[i];
        const int C = u.Cout;
        const int up = c.voc_up_rates[i];
        const int lout = len * up;
        const int lpad_in = lpad;
        lpad = atoms_lpad(lout);
        for (__half* b : {XA, TA, RA[0], RA[1]}) launch_atoms_zero_pads(b, nb * C / 8, lpad, lout, st);
        launch_convT_tc(PA, u.blob.p, u.plan, u.b.p, cb + cbias_off[i + 1], X, XA, u.Cin, C, len, lpad_in, lpad, up, 0.1f, nb,
                        cbias_stride, st);
        len = lout;
        { char nm[16]; snprintf(nm, sizeof(nm), "up%d", i); dump(nm, X, (size_t)C * len); }
        const bool more = (i + 1 < c.voc_n_up);
        if (more) launch_atoms_zero_pads(PA, nb * C / 8, lpad, len, st);     // PA is re-shaped for the next stage's input
        for (int j = 0; j < nk; ++j) {
            const RB& rb = *rbs[i * nk + j];
            const __half* in16 = XA;
            const float* resid = X;
            for (int t = 0; t < 3; ++t) {
                conv1d_tc(*rb.c1[t], in16, nullptr, nullptr, nullptr, TA, len, lpad, c.voc_rb_dilations[t], CONV_STORE, nb);
                if (t < 2) {
                    conv1d_tc(*rb.c2[t], TA, nullptr, resid, R[t], RA[t], len, lpad, 1, CONV_STORE, nb);
                    in16 = RA[t]; resid = R[t];
                } else {
                    // MRF sum; the last resblock also emits lrelu(sum / nk) as the next transposed conv's operand
                    const bool emit = more && (j == nk - 1);
                    conv1d_tc(*rb.c2[t], TA, nullptr, resid, ZS, emit ? PA : nullptr, len, lpad, 1,
                              j == 0 ? CONV_STORE : CONV_ACCUM, nb, 1.0f / (float)nk);
                }
            }
        }
        { char nm[16]; snprintf(nm, sizeof(nm), "mrf%d", i); dump(nm, ZS, (size_t)C * len); }
    }
    const float* cur = ZS;
    const float in_scale = 1.0f / (float)nk;
    launch_conv_post(cur, conv_post_w.p, wav_dev_out, post_cin, len, 7, in_scale, 0.01f, nb, st);
    *n_out = len;
}

// ================================================================================================
// scheduler
// ================================================================================================
float* Engine::pinned_get(size_t n, size_t* cap) {
    std::lock_guard<std::mutex> lk(pin_mu);
    for (size_t i = 0; i < pinned_pool.size(); ++i)
        if (pinned_pool[i].second >= n) {
            float* p = pinned_pool[i].first; *cap = pinned_pool[i].second;
            pinned_pool.erase(pinned_pool.begin() + i);
            return p;
        }
    float* p = nullptr;
    CUDA_CHECK(cudaMallocHost(&p, n * sizeof(float)));
    *cap = n;
    return p;
}
float* Engine::dev_get(size_t n, size_t* cap) {
    std::lock_guard<std::mutex> lk(pin_mu);
    size_t best = dev_pool.size();
    for (size_t i = 0; i < dev_pool.size(); ++i)
        if (dev_pool[i].second >= n && (best == dev_pool.size() || dev_pool[i].second < dev_pool[best].second)) best = i;
    if (best != dev_pool.size()) {
        float* p = dev_pool[best].first; *cap = dev_pool[best].second;
        dev_pool.erase(dev_pool.begin() + best);
        return p;
    }
    float* p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, n * sizeof(float)));
    *cap = n;
    return p;
}
void Engine::dev_put(float* p, size_t cap) { std::lock_guard<std::mutex> lk(pin_mu); dev_pool.emplace_back(p, cap); }

void Engine::pinned_put(float* p, size_t cap) { std::lock_guard<std::mutex> lk(pin_mu); pinned_pool.emplace_back(p, cap); }

void Engine::submit(uint64_t id, const int32_t* text, int n_text, int speaker, const xtts_sampling& sp) {
    if (n_text <= 0 || n_text > cfg.max_text_tokens + 2) throw std::runtime_error("n_text out of range (1..max_text_tokens+2)");
    if (speaker < 0 || speaker >= S) throw std::runtime_error("speaker slot out of range");
    std::shared_ptr<Sequence> s(new Sequence());
    s->id = id; s->text_ids.assign(text, text + n_text); s->speaker = speaker; s->sp = sp; s->t_submit = now_s();
    if (id & kPartialBit) throw std::runtime_error("seq_id must be < 2^63");
    if (sp.early_tokens > 0 && sp.vocode) s->early_tokens = sp.early_tokens;
    require_finalized();
    if (!spk_valid[speaker]) throw std::runtime_error("speaker slot not set");
    {
        std::lock_guard<std::mutex> lk(q_mu);
        pending.push_back(s);
        ++inflight;
    }
    cv_work.notify_all();
}

// tokens + latent snapshot of one finished sequence; frees its slot
void Engine::finish_sequence(std::shared_ptr<Sequence> s) {
    int n = 0;
    d_n_gen.download(&n, 1, st, s->slot);
    CUDA_CHECK(cudaStreamSynchronize(st));
    n = std::min(n, CAP);
    s->tokens.resize(n);
    d_tokens.download(s->tokens.data(), n, st, (size_t)s->slot * CAP);
    // latents copy (device) so the slot can be recycled immediately
    s->lat_dev = dev_get((size_t)std::max(1, n) * H, &s->lat_dev_cap);
    CUDA_CHECK(cudaMemcpyAsync(s->lat_dev, d_latents.p + (size_t)s->slot * CAP * H, (size_t)n * H * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    st_tokens += n;
    release_slot(*s);
}

// vocode a group of finished sequences that have the same token count, VB at a time, then retire them
void Engine::finish_group(std::vector<std::shared_ptr<Sequence>>& grp) {
    const double t0 = now_s();
    for (size_t b0 = 0; b0 < grp.size(); b0 += VB) {
        const int nb = (int)std::min<size_t>(VB, grp.size() - b0);
        const int n = (int)grp[b0]->tokens.size();
        std::vector<int> spk(nb);
        for (int i = 0; i < nb; ++i) {
            spk[i] = grp[b0 + i]->speaker;
            CUDA_CHECK(cudaMemcpyAsync(vlat.p + (size_t)i * n * H, grp[b0 + i]->lat_dev, (size_t)n * H * sizeof(float),
                                       cudaMemcpyDeviceToDevice, st));
        }
        int ns = 0;
        run_vocoder(vlat.p, n, spk.data(), nb, vwav.p, &ns, nullptr, nullptr, 0);
        for (int i = 0; i < nb; ++i) {
            auto& s = grp[b0 + i];
            const int skip = std::min(s->early_samples, ns);       // 0 unless a partial first-audio piece went out already
            const int nrem = ns - skip;
            const float* src = vwav.p + (size_t)i * ns + skip;
            s->n_samples = nrem;
            if (d2h_wav) {
                s->wav_host = pinned_get(std::max(1, nrem), &s->wav_cap);
                if (nrem > 0) CUDA_CHECK(cudaMemcpyAsync(s->wav_host, src, (size_t)nrem * sizeof(float), cudaMemcpyDeviceToHost, st));
            } else {
                s->wav_dev = dev_get(std::max(1, nrem), &s->wav_dev_cap);
                if (nrem > 0) CUDA_CHECK(cudaMemcpyAsync(s->wav_dev, src, (size_t)nrem * sizeof(float), cudaMemcpyDeviceToDevice, st));
            }
            st_samples += nrem;
        }
        CUDA_CHECK(cudaStreamSynchronize(st));
        const double t = now_s();
        for (int i = 0; i < nb; ++i) { grp[b0 + i]->t_done = t; retire(grp[b0 + i]); }
    }
    st_voc_ms += (now_s() - t0) * 1e3;
}

// samples the vocoder produces for T latent frames (the two interpolations of HifiDecoder.forward, then the upsampling)
int Engine::samples_for(int T) const {
    const double s1 = (double)cfg.code_stride / (double)cfg.output_hop_length;
    const double s2 = (double)cfg.output_sample_rate / (double)cfg.input_sample_rate;
    const int T1 = (int)std::floor((double)T * s1);
    int len = cfg.output_sample_rate != cfg.input_sample_rate ? (int)std::floor((double)T1 * s2) : T1;
    for (int i = 0; i < cfg.voc_n_up; ++i) len *= cfg.voc_up_rates[i];
    return len;
}

// First audio early (xtts_sampling.early_tokens; not in the reference, which returns whole chunks):
// every sequence of `grp` is still decoding and has early_tokens + kEarlyLookahead latent frames.  Vocode that prefix
// straight from the latent ring and deliver the samples of the first early_tokens frames as a PARTIAL result (status 1);
// the lookahead frames cover the vocoder's receptive field, so these samples equal the ones the whole chunk will give.
void Engine::emit_early(std::vector<std::shared_ptr<Sequence>>& grp) {
    const double t0 = now_s();
    std::map<int, std::vector<std::shared_ptr<Sequence>>> by_len;
    for (auto& s : grp) by_len[s->early_tokens].push_back(s);
    for (auto& kv : by_len) {
        const int n_early = kv.first, T = n_early + kEarlyLookahead;
        const int keep = samples_for(n_early);
        auto& v = kv.second;
        for (size_t b0 = 0; b0 < v.size(); b0 += VB) {
            const int nb = (int)std::min<size_t>(VB, v.size() - b0);
            std::vector<int> spk(nb);
            for (int i = 0; i < nb; ++i) {
                spk[i] = v[b0 + i]->speaker;
                CUDA_CHECK(cudaMemcpyAsync(vlat.p + (size_t)i * T * H, d_latents.p + (size_t)v[b0 + i]->slot * CAP * H,
                                           (size_t)T * H * sizeof(float), cudaMemcpyDeviceToDevice, st));
            }
            int ns = 0;
            run_vocoder(vlat.p, T, spk.data(), nb, vwav.p, &ns, nullptr, nullptr, 0);
            if (keep <= 0 || keep > ns) throw std::runtime_error("early emit: sample count out of range");
            std::vector<std::shared_ptr<Sequence>> parts;
            for (int i = 0; i < nb; ++i) {
                auto& s = v[b0 + i];
                std::shared_ptr<Sequence> p(new Sequence());
                p->id = s->id; p->speaker = s->speaker; p->status = 1; p->n_prompt = s->n_prompt;
                p->t_submit = s->t_submit; p->t_first = s->t_first;
                p->tokens.resize(n_early);
                d_tokens.download(p->tokens.data(), n_early, st, (size_t)s->slot * CAP);
                p->n_samples = keep;
                if (d2h_wav) {
                    p->wav_host = pinned_get(keep, &p->wav_cap);
                    CUDA_CHECK(cudaMemcpyAsync(p->wav_host, vwav.p + (size_t)i * ns, (size_t)keep * sizeof(float), cudaMemcpyDeviceToHost, st));
                } else {
                    p->wav_dev = dev_get(keep, &p->wav_dev_cap);
                    CUDA_CHECK(cudaMemcpyAsync(p->wav_dev, vwav.p + (size_t)i * ns, (size_t)keep * sizeof(float), cudaMemcpyDeviceToDevice, st));
                }
                st_samples += keep;
                parts.push_back(p);
            }
            CUDA_CHECK(cudaStreamSynchronize(st));
            const double t = now_s();
            for (int i = 0; i < nb; ++i) {
                v[b0 + i]->early_done = true; v[b0 + i]->early_samples = keep;
                parts[i]->t_done = t;
                std::lock_guard<std::mutex> lk(q_mu);          // a partial piece does not end the sequence: inflight unchanged
                done_q.push_back(parts[i]);
                done_map[parts[i]->id | kPartialBit] = parts[i];
            }
            cv_done.notify_all();
        }
    }
    st_voc_ms += (now_s() - t0) * 1e3;
}

void Engine::retire(std::shared_ptr<Sequence> s) {
    {
        std::lock_guard<std::mutex> lk(q_mu);
        done_q.push_back(s);
        done_map[s->id] = s;
        --inflight;
    }
    cv_done.notify_all();
}

void Engine::loop() {
    cudaSetDevice(cfg.device);
    while (true) {
        {
            std::unique_lock<std::mutex> q(q_mu);
            cv_work.wait(q, [&] {
                return stop.load() || !pending.empty() || (!waiting.empty() && !hold_admission.load()) || !running.empty();
            });
            if (stop.load()) break;
            while (!pending.empty()) {
                // stable priority insert: lower `priority` (chunk index) first, FIFO among equals
                auto s = pending.front(); pending.pop_front();
                auto it = waiting.end();
                while (it != waiting.begin() && (*(it - 1))->sp.priority > s->sp.priority) --it;
                waiting.insert(it, s);
            }
        }
        std::lock_guard<std::mutex> lk(mu);
        auto fail = [&](std::shared_ptr<Sequence>& s, int code, const char* what) {
            s->status = code; set_error(what);
            release_slot(*s); s->t_done = now_s();
            retire(s);
        };
        std::vector<std::shared_ptr<Sequence>> fresh_sp;      // admitted this iteration (outside the try: see the catch)
        try {
            // ---- admission (continuous batching): fill free slots, whole prompts, within the row budget
            std::vector<Sequence*> fresh;
            int rows = 0;
            while (!hold_admission.load() && !waiting.empty() && !free_slots.empty()) {
                auto s = waiting.front();
                const int p = cfg.n_cond_latents + (int)s->text_ids.size() + 1;
                if (!fresh.empty() && rows + p > prefill_rows_cap) break;
                waiting.pop_front();
                s->slot = free_slots.back(); free_slots.pop_back();
                try {
                    if (!spk_valid[s->speaker]) throw std::runtime_error("speaker slot not set");
                    init_slot(*s, nullptr, 0);
                } catch (const std::exception& ex) {
                    fail(s, XTTS_ERR_STATE, ex.what());
                    continue;
                }
                rows += p;
                fresh.push_back(s.get()); fresh_sp.push_back(s);
            }
            const double t0 = now_s();
            if (!fresh.empty()) {
                prefill(fresh);
                for (auto& s : fresh_sp) running.push_back(s);
                fresh_sp.clear();
                // a sequence may already be finished after its first token (max_tokens == 1 / instant stop)
                d_finished.download(h_finished, NSLOT, st);
                CUDA_CHECK(cudaStreamSynchronize(st));
            }
            if (!running.empty()) {
                std::vector<int> active;
                double ctx_sum = 0;
                for (auto& s : running) if (!h_finished[s->slot]) { active.push_back(s->slot); ctx_sum += s->n_prompt + s->steps + 1; ++s->steps; }
                if (!active.empty()) {
                    decode_ctx_sum = ctx_sum;
                    decode_step(active);
                    d_finished.download(h_finished, NSLOT, st);
                    CUDA_CHECK(cudaStreamSynchronize(st));
                    if (last_prof) { g_prof.collect_graph(*last_prof, decode_ctx_sum); last_prof = nullptr; }
                }
            }
            st_gpt_ms += (now_s() - t0) * 1e3;
            // ---- first audio early (off unless a chunk asked for it): sequences that just reached their prefix length
            {
                std::vector<std::shared_ptr<Sequence>> early;
                for (auto& s : running)
                    if (s->early_tokens > 0 && !s->early_done && !h_finished[s->slot] &&
                        s->steps + 1 >= s->early_tokens + kEarlyLookahead && s->early_tokens + kEarlyLookahead <= voc_max_T)
                        early.push_back(s);
                if (!early.empty()) emit_early(early);
            }
            // ---- retire finished sequences: vocode, D2H, completion queue
            std::vector<std::shared_ptr<Sequence>> keep, fin;
            for (auto& s : running) (h_finished[s->slot] ? fin : keep).push_back(s);
            running.swap(keep);
            std::map<int, std::vector<std::shared_ptr<Sequence>>> groups;     // token count -> sequences to vocode
            for (auto& s : fin) {
                try {
                    finish_sequence(s);
                    if (s->sp.vocode) groups[(int)s->tokens.size()].push_back(s);
                    else { s->t_done = now_s(); retire(s); }
                } catch (const std::exception& ex) { fail(s, XTTS_ERR_CUDA, ex.what()); }
            }
            for (auto& kv : groups) {
                try { finish_group(kv.second); }
                catch (const std::exception& ex) {
                    for (auto& s : kv.second) if (s->t_done == 0) fail(s, XTTS_ERR_CUDA, ex.what());
                }
            }
        } catch (const std::exception& ex) {
            // a failure inside a batched step fails every sequence that was part of it — including the ones admitted
            // in this iteration whose prefill threw before they reached `running`
            for (auto& s : fresh_sp) fail(s, XTTS_ERR_CUDA, ex.what());
            for (auto& s : running) fail(s, XTTS_ERR_CUDA, ex.what());
            running.clear();
        }
    }
}

int Engine::poll(xtts_result* out, int timeout_ms) {
    std::unique_lock<std::mutex> lk(q_mu);
    if (!cv_done.wait_for(lk, std::chrono::milliseconds(std::max(0, timeout_ms)), [&] { return !done_q.empty(); })) return 0;
    auto s = done_q.front(); done_q.pop_front();
    out->seq_id = s->id; out->status = s->status; out->n_tokens = (int)s->tokens.size(); out->n_samples = s->n_samples;
    out->n_prompt_rows = s->n_prompt; out->t_submit = s->t_submit; out->t_first_token = s->t_first; out->t_done = s->t_done;
    return 1;
}

void Engine::fetch(uint64_t id, int32_t* tokens, float* wav, float* latents) {
    std::shared_ptr<Sequence> s;
    {
        std::lock_guard<std::mutex> lk(q_mu);
        auto it = done_map.find(id | kPartialBit);          // an unfetched first-audio piece of this id is older than its final
        if (it == done_map.end()) it = done_map.find(id);
        if (it == done_map.end()) throw std::runtime_error("fetch: unknown or unfinished sequence id");
        s = it->second;
        done_map.erase(it);
        for (auto q = done_q.begin(); q != done_q.end(); ++q) if (*q == s) { done_q.erase(q); break; }
    }
    if (tokens) std::memcpy(tokens, s->tokens.data(), s->tokens.size() * sizeof(int32_t));
    if (wav && s->n_samples > 0 && s->wav_host) std::memcpy(wav, s->wav_host, (size_t)s->n_samples * sizeof(float));
    const bool dev_wav = wav && s->n_samples > 0 && !s->wav_host && s->wav_dev;
    const bool dev_lat = latents && s->lat_dev;
    if (dev_wav || dev_lat) {
        std::lock_guard<std::mutex> lk(mu);          // device copies go through the engine stream
        CUDA_CHECK(cudaSetDevice(cfg.device));
        if (dev_wav) CUDA_CHECK(cudaMemcpyAsync(wav, s->wav_dev, (size_t)s->n_samples * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (dev_lat) CUDA_CHECK(cudaMemcpyAsync(latents, s->lat_dev, s->tokens.size() * (size_t)H * sizeof(float), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
    }
    if (s->wav_host) { pinned_put(s->wav_host, s->wav_cap); s->wav_host = nullptr; }
    if (s->wav_dev) { dev_put(s->wav_dev, s->wav_dev_cap); s->wav_dev = nullptr; }
    if (s->lat_dev) { dev_put(s->lat_dev, s->lat_dev_cap); s->lat_dev = nullptr; }
}

void Engine::set_option(const std::string& k, int64_t v) {
    if (k == "hold_admission") {       // no GPU state involved: must not wait for a running step
        hold_admission.store(v != 0);
        { std::lock_guard<std::mutex> q(q_mu); }
        cv_work.notify_all();
        return;
    }
    std::lock_guard<std::mutex> lk(mu);
    if (k == "d2h_wav") d2h_wav = v != 0;
    else if (k == "tc_vocoder") use_tc_vocoder = v != 0;
    else if (k == "conv_epi_groups") g_conv_epi_groups = v >= 2 ? 2 : 1;
    else if (k == "attn_ctas_per_sm" || k == "gemm_bn") {
        if (k == "gemm_bn") { if (v != 0 && v != 32 && v != 64 && v != 128) throw std::runtime_error("gemm_bn: 0, 32, 64 or 128"); g_gemm_decode_bn = (int)v; }
        else g_attn_ctas_per_sm = (int)v;                        // < 0: absolute grid size (tests)
        drop_graphs();
    }
    else if (k == "cuda_graphs") use_graphs = v != 0;
    else if (k == "pdl") { use_pdl = v != 0; drop_graphs(); }
    else if (k == "splitk") { use_splitk = v != 0; drop_graphs(); }
    else if (k == "decode_chain") { use_chain = v != 0; drop_graphs(); }
    else if (k == "microbatches" || k == "microbatch_min_rows") {
        if (k == "microbatches") n_micro = std::max<int>(1, std::min<int64_t>(v, kMaxMicro)); else micro_min_rows = (int)std::max<int64_t>(2, v);
        drop_graphs();
    }
    else if (k == "profile") { CUDA_CHECK(cudaSetDevice(cfg.device)); CUDA_CHECK(cudaStreamSynchronize(st)); g_prof.reset(); g_prof.enabled = v != 0; }
    else if (k == "reset_stats") {
        st_decode_steps = st_prefill_rows = st_tokens = st_samples = 0; st_gpt_ms = st_voc_ms = st_cond_ms = 0;
        launch_base = g_launch_count;
    } else throw std::runtime_error("unknown option: " + k);
}

void Engine::get_stats(xtts_stats* s) {
    std::lock_guard<std::mutex> lk(mu);
    s->kernel_launches = g_launch_count - launch_base; s->decode_steps = st_decode_steps; s->prefill_rows = st_prefill_rows;
    s->tokens_generated = st_tokens; s->samples_generated = st_samples; s->gpt_ms = st_gpt_ms; s->vocoder_ms = st_voc_ms;
    s->cond_ms = st_cond_ms; s->hbm_bytes_weights = weight_bytes;
}

void Engine::kernel_profile(xtts_kernel_profile* out) {
    std::lock_guard<std::mutex> lk(mu);
    CUDA_CHECK(cudaSetDevice(cfg.device));
    CUDA_CHECK(cudaStreamSynchronize(st));
    g_prof.collect();
    std::memset(out, 0, sizeof(*out));
    out->n = KF_COUNT;
    for (int i = 0; i < KF_COUNT && i < 16; ++i) {
        std::snprintf(out->name[i], sizeof(out->name[i]), "%s", kernel_family_name(i));
        out->ms[i] = g_prof.ms[i]; out->flops[i] = g_prof.flops[i]; out->bytes[i] = g_prof.bytes[i];
        out->launches[i] = g_prof.launches[i];
    }
}

// Stopwatch on the engine stream.  Every other stream the engine uses (decode branches) forks from and joins back into
// `st` inside a step, so an event recorded on `st` behind a step completes after all of that step's device work.
void Engine::device_timer(int op, double* ms) {
    std::lock_guard<std::mutex> lk(mu);
    CUDA_CHECK(cudaSetDevice(cfg.device));
    if (!ev_t0) { CUDA_CHECK(cudaEventCreate(&ev_t0)); CUDA_CHECK(cudaEventCreate(&ev_t1)); }
    if (op == 0) {
        CUDA_CHECK(cudaEventRecord(ev_t0, st));
        timer_armed = true;
    } else if (op == 1) {
        if (!timer_armed) throw std::runtime_error("device_timer: stop without start");
        CUDA_CHECK(cudaEventRecord(ev_t1, st));
        CUDA_CHECK(cudaEventSynchronize(ev_t1));
        float t = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&t, ev_t0, ev_t1));
        if (ms) *ms = (double)t;
    } else throw std::runtime_error("device_timer: op must be 0 (start) or 1 (stop)");
}

void Engine::sync_idle() {
    std::unique_lock<std::mutex> lk(q_mu);
    cv_done.wait(lk, [&] { return inflight == 0; });
}

// ================================================================================================
// synchronous single-stage entry points
// ================================================================================================
void Engine::vocode_sync(const float* latents, int T, int speaker, float* wav, int* n_out, const char* stage,
                         float* stage_out, int64_t stage_cap) {
    std::lock_guard<std::mutex> lk(mu);
    require_finalized();
    CUDA_CHECK(cudaSetDevice(cfg.device));
    DBuf<float> lat; lat.alloc((size_t)T * cfg.voc_in_dim);
    lat.upload(latents, (size_t)T * cfg.voc_in_dim, st);
    int ns = 0;
    run_vocoder(lat.p, T, &speaker, 1, vwav.p, &ns, stage, stage_out, stage_cap);
    if (wav) vwav.download(wav, ns, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (n_out) *n_out = ns;
}

void Engine::gpt_prefill_sync(const int32_t* text, int n_text, int speaker, const int32_t* audio, int n_audio,
                              float* hidden_out, float* logits_out, float* latents_out) {
    std::lock_guard<std::mutex> lk(mu);
    require_finalized();
    if (!running.empty() || !waiting.empty()) throw std::runtime_error("debug entry points need an idle engine");
    CUDA_CHECK(cudaSetDevice(cfg.device));
    Sequence s; s.text_ids.assign(text, text + n_text); s.speaker = speaker; s.slot = B;
    s.sp.max_tokens = CAP; s.sp.stop_token = cfg.stop_audio_token; s.sp.repetition_penalty = 1.f; s.sp.temperature = 0.f;
    init_slot(s, nullptr, 0);
    const int n = std::max(1, n_audio);
    std::vector<std::vector<int32_t>> aud(1);
    if (n_audio > 1) aud[0].assign(audio, audio + n_audio - 1);       // rows fed by t_1..t_{n-1}
    std::vector<int> last_rows; int max_nq = 0;
    std::vector<Sequence*> seqs{&s};
    const int M = build_prefill(seqs, aud, last_rows, max_nq);
    launch_build_rows(d_rows.p, M, tables(), wX.p, st);
    layers_forward(M, true, 1, max_nq);
    if (hidden_out) {
        // ln_f of every row
        launch_layernorm<float>(wX.p, lnfw.p, lnfb.p, wQKV.p, M, H, cfg.ln_eps, st);
        CUDA_CHECK(cudaMemcpyAsync(hidden_out, wQKV.p, (size_t)M * H * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    std::vector<int> ridx(n), slots(n, B), lpos(n);
    for (int i = 0; i < n; ++i) { ridx[i] = M - n + i; lpos[i] = i; }
    d_rowidx.upload(ridx.data(), n, st); d_active.upload(slots.data(), 1, st);
    DBuf<int> dslots; dslots.alloc(n); dslots.upload(slots.data(), n, st);
    d_lat_pos.upload(lpos.data(), n, st);
    head_and_sample(n, d_rowidx.p, dslots.p, d_lat_pos.p, 0, false);
    if (logits_out)
        CUDA_CHECK(cudaMemcpy2DAsync(logits_out, (size_t)V * sizeof(float), wLOG.p, (size_t)Vpad * sizeof(float),
                                     (size_t)V * sizeof(float), n, cudaMemcpyDeviceToHost, st));
    if (latents_out) d_latents.download(latents_out, (size_t)n * H, st, (size_t)B * CAP * H);
    CUDA_CHECK(cudaStreamSynchronize(st));
    release_slot(s);
}

void Engine::gpt_teacher_forced_sync(const int32_t* text, int n_text, int speaker, const int32_t* forced, int n,
                                     const xtts_sampling& sp, float* logits_out, float* latents_out, int32_t* sampled_out) {
    std::lock_guard<std::mutex> lk(mu);
    require_finalized();
    if (!running.empty() || !waiting.empty()) throw std::runtime_error("debug entry points need an idle engine");
    if (n < 1 || n > CAP) throw std::runtime_error("teacher_forced: n out of range");
    CUDA_CHECK(cudaSetDevice(cfg.device));
    Sequence s; s.text_ids.assign(text, text + n_text); s.speaker = speaker; s.slot = B; s.sp = sp;
    s.sp.max_tokens = n;
    use_forced = forced != nullptr;
    try {
        init_slot(s, forced, n);
        std::vector<Sequence*> seqs{&s};
        prefill(seqs);                                   // samples token 1 (forced -> t_1)
        if (logits_out)
            CUDA_CHECK(cudaMemcpyAsync(logits_out, wLOG.p, (size_t)V * sizeof(float), cudaMemcpyDeviceToHost, st));
        std::vector<int> active{B};
        for (int k = 1; k < n; ++k) {
            decode_step(active);
            if (logits_out)
                CUDA_CHECK(cudaMemcpyAsync(logits_out + (size_t)k * V, wLOG.p, (size_t)V * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        if (latents_out) d_latents.download(latents_out, (size_t)n * H, st, (size_t)B * CAP * H);
        if (sampled_out) d_sampled.download(sampled_out, n, st, (size_t)B * CAP);
        CUDA_CHECK(cudaStreamSynchronize(st));
    } catch (...) {
        use_forced = false; release_slot(s);
        throw;
    }
    use_forced = false;
    release_slot(s);
}

void Engine::debug_gemm(int mode, const float* A, const float* W, const float* bias, const float* resid, float* out, int M,
                        int N, int K, int gelu, int iters, float* ms) {
    std::lock_guard<std::mutex> lk(mu);
    CUDA_CHECK(cudaSetDevice(cfg.device));
    DBuf<float> dA, dW, db, dr, dout;
    dA.alloc((size_t)M * K); dW.alloc((size_t)N * K); dout.alloc((size_t)M * N);
    dA.upload(A, (size_t)M * K, st); dW.upload(W, (size_t)N * K, st);
    if (bias) { db.alloc(N); db.upload(bias, N, st); }
    if (resid) { dr.alloc((size_t)M * N); dr.upload(resid, (size_t)M * N, st); }
    const int flags = (gelu ? GEMM_GELU : 0) | (resid ? GEMM_RESID : 0);
    DBuf<__nv_bfloat16> hA, hW;
    if (mode == 1) {
        std::string err;
        if (!gemm_tc_init(&err)) throw std::runtime_error(err);
        hA.alloc((size_t)M * K); hW.alloc((size_t)N * K);
        launch_f32_to_bf16(dA.p, hA.p, (size_t)M * K, st); launch_f32_to_bf16(dW.p, hW.p, (size_t)N * K, st);
    }
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
    auto run = [&] {
        if (mode == 1) launch_gemm_bf16_tc(hA.p, hW.p, db.p, dr.p, dout.p, M, N, K, flags, st);
        else launch_gemm_f32(dA.p, dW.p, db.p, dr.p, dout.p, M, N, K, flags, st);
    };
    run();
    CUDA_CHECK(cudaStreamSynchronize(st));
    dout.download(out, (size_t)M * N, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (iters > 0) {
        CUDA_CHECK(cudaEventRecord(e0, st));
        for (int i = 0; i < iters; ++i) run();
        CUDA_CHECK(cudaEventRecord(e1, st));
        CUDA_CHECK(cudaEventSynchronize(e1));
        float t = 0; CUDA_CHECK(cudaEventElapsedTime(&t, e0, e1));
        if (ms) *ms = t / iters;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

void Engine::debug_sample(const float* logits, const uint8_t* seen, int Bn, int Vn, const xtts_sampling& sp, int step,
                          int32_t* out) {
    std::lock_guard<std::mutex> lk(mu);
    if (!running.empty() || !waiting.empty()) throw std::runtime_error("debug entry points need an idle engine");
    if (Bn < 1 || Bn > B || Vn != V) throw std::runtime_error("debug_sample: bad batch or vocabulary size");
    CUDA_CHECK(cudaSetDevice(cfg.device));
    std::vector<float> lg((size_t)Bn * Vpad, 0.f);
    for (int b = 0; b < Bn; ++b) std::memcpy(&lg[(size_t)b * Vpad], logits + (size_t)b * V, V * sizeof(float));
    wLOG.upload(lg.data(), lg.size(), st);
    std::vector<int> slots(Bn), ngen(Bn, step), zeros(Bn, 0), tk(Bn, sp.top_k), mt(Bn, CAP), stp(Bn, sp.stop_token), ss(Bn);
    std::vector<float> T(Bn, sp.temperature), tp(Bn, sp.top_p), pen(Bn, sp.repetition_penalty);
    std::vector<unsigned long long> seed(Bn, sp.seed);
    std::vector<unsigned> sb((size_t)Bn * SEENW, 0u);
    for (int b = 0; b < Bn; ++b) {
        slots[b] = b; ss[b] = sp.seq_seed + b;
        if (seen) for (int v = 0; v < V; ++v) if (seen[(size_t)b * V + v]) sb[(size_t)b * SEENW + (v >> 5)] |= 1u << (v & 31);
    }
    d_active.upload(slots.data(), Bn, st); d_n_gen.upload(ngen.data(), Bn, st); d_finished.upload(zeros.data(), Bn, st);
    d_top_k.upload(tk.data(), Bn, st); d_max_tokens.upload(mt.data(), Bn, st); d_stop.upload(stp.data(), Bn, st);
    d_seq_seed.upload(ss.data(), Bn, st); d_temp.upload(T.data(), Bn, st); d_top_p.upload(tp.data(), Bn, st);
    d_pen.upload(pen.data(), Bn, st); d_seed.upload(seed.data(), Bn, st); d_seen.upload(sb.data(), sb.size(), st);
    launch_sample(wLOG.p, Vpad, d_active.p, Bn, V, sample_state(), 0, st);
    std::vector<int> res(Bn);
    d_last_tok.download(res.data(), Bn, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    for (int b = 0; b < Bn; ++b) out[b] = res[b];
    d_n_gen.upload(zeros.data(), Bn, st); d_finished.upload(zeros.data(), Bn, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
}

}  // namespace xtts

// ================================================================================================
// C ABI
// ================================================================================================
using xtts::Engine;
struct xtts_engine { Engine* impl; };

#define XTTS_TRY(body)                                                  \
    try { body; return XTTS_OK; }                                       \
    catch (const xtts::CudaError& ex) { xtts::set_error(ex.what()); return XTTS_ERR_CUDA; } \
    catch (const std::exception& ex) { xtts::set_error(ex.what()); return XTTS_ERR_INVALID; } \
    catch (...) { xtts::set_error("unknown error"); return XTTS_ERR_INVALID; }

extern "C" {

const char* xtts_last_error(void) {
    if (!xtts::t_last_error.empty()) return xtts::t_last_error.c_str();
    std::lock_guard<std::mutex> lk(xtts::g_err_mu);
    xtts::t_last_error = xtts::g_last_error;
    return xtts::t_last_error.c_str();
}
const char* xtts_version(void) { return "libxtts_b200 0.1 (sm_100a)"; }

int xtts_create(const xtts_config* cfg, xtts_engine** out) {
    if (!cfg || !out) { xtts::set_error("null argument"); return XTTS_ERR_INVALID; }
    XTTS_TRY({ Engine* e = new Engine(*cfg); *out = new xtts_engine{e}; })
}
int xtts_destroy(xtts_engine* e) {
    if (!e) return XTTS_OK;
    XTTS_TRY({ delete e->impl; delete e; })
}
int xtts_load_weight(xtts_engine* e, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
    XTTS_TRY(e->impl->load_weight(name, data, shape, ndim))
}
int xtts_finalize_weights(xtts_engine* e) { XTTS_TRY(e->impl->finalize_weights()) }
int xtts_set_speaker(xtts_engine* e, int32_t slot, const float* cond, const float* g) { XTTS_TRY(e->impl->set_speaker(slot, cond, g)) }
int xtts_get_speaker(xtts_engine* e, int32_t slot, float* cond, float* g) { XTTS_TRY(e->impl->get_speaker(slot, cond, g)) }
int xtts_condition(xtts_engine* e, int32_t slot, const float* wav22k, int64_t n22, const float* wav16k, int64_t n16,
                   int32_t gpt_cond_len_s, int32_t gpt_cond_chunk_len_s) {
    XTTS_TRY(e->impl->condition(slot, wav22k, n22, wav16k, n16, gpt_cond_len_s, gpt_cond_chunk_len_s))
}
int xtts_submit(xtts_engine* e, uint64_t seq_id, const int32_t* text_ids, int32_t n_text, int32_t speaker_slot,
                const xtts_sampling* sp) {
    XTTS_TRY(e->impl->submit(seq_id, text_ids, n_text, speaker_slot, *sp))
}
int xtts_poll(xtts_engine* e, xtts_result* out, int32_t timeout_ms) {
    try { return e->impl->poll(out, timeout_ms); }
    catch (const std::exception& ex) { xtts::set_error(ex.what()); return XTTS_ERR_INVALID; }
}
int xtts_fetch(xtts_engine* e, uint64_t seq_id, int32_t* tokens, float* wav, float* latents) {
    XTTS_TRY(e->impl->fetch(seq_id, tokens, wav, latents))
}
int xtts_set_option(xtts_engine* e, const char* key, int64_t value) { XTTS_TRY(e->impl->set_option(key, value)) }
int xtts_get_stats(xtts_engine* e, xtts_stats* out) { XTTS_TRY(e->impl->get_stats(out)) }
int xtts_sync(xtts_engine* e) { XTTS_TRY(e->impl->sync_idle()) }
int xtts_get_kernel_profile(xtts_engine* e, xtts_kernel_profile* out) { XTTS_TRY(e->impl->kernel_profile(out)) }
int xtts_device_timer(xtts_engine* e, int32_t op, double* ms) { XTTS_TRY(e->impl->device_timer(op, ms)) }
int xtts_vocode(xtts_engine* e, const float* latents, int32_t T, int32_t speaker_slot, float* wav, int32_t* n_out,
                const char* stage, float* stage_out, int64_t stage_cap) {
    XTTS_TRY(e->impl->vocode_sync(latents, T, speaker_slot, wav, n_out, stage, stage_out, stage_cap))
}
int xtts_gpt_prefill(xtts_engine* e, const int32_t* text_ids, int32_t n_text, int32_t speaker_slot,
                     const int32_t* audio_tokens, int32_t n_audio, float* hidden_out, float* logits_out, float* latents_out) {
    XTTS_TRY(e->impl->gpt_prefill_sync(text_ids, n_text, speaker_slot, audio_tokens, n_audio, hidden_out, logits_out, latents_out))
}
int xtts_gpt_teacher_forced(xtts_engine* e, const int32_t* text_ids, int32_t n_text, int32_t speaker_slot,
                            const int32_t* forced_tokens, int32_t n, const xtts_sampling* sp, float* logits_out,
                            float* latents_out, int32_t* sampled_out) {
    XTTS_TRY(e->impl->gpt_teacher_forced_sync(text_ids, n_text, speaker_slot, forced_tokens, n, *sp, logits_out, latents_out, sampled_out))
}
int xtts_debug_gemm(xtts_engine* e, int32_t mode, const float* A, const float* W, const float* bias, const float* resid,
                    float* out, int32_t M, int32_t N, int32_t K, int32_t gelu, int32_t iters, float* ms_per_iter) {
    XTTS_TRY(e->impl->debug_gemm(mode, A, W, bias, resid, out, M, N, K, gelu, iters, ms_per_iter))
}
int xtts_debug_sample(xtts_engine* e, const float* logits, const uint8_t* seen, int32_t B, int32_t V,
                      const xtts_sampling* sp, int32_t step, int32_t* out_tokens) {
    XTTS_TRY(e->impl->debug_sample(logits, seen, B, V, *sp, step, out_tokens))
}

}  // extern "C"
