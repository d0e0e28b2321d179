// Speaker conditioning on the GPU: reference wav -> (GPT conditioning latents [32,H], d-vector [proj]).
// Replaces get_conditioning_latents (XTTSv2.py:409-468): wav_to_mel_cloning -> ConditioningEncoder ->
// PerceiverResampler (per 4 s piece, averaged) and ResNetSpeakerEncoder.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/xtts_b200.h"
#include "kernels.h"

namespace xtts {

struct HostTensorView {
    const float* data;
    std::vector<int64_t> shape;
    size_t numel() const { size_t n = 1; for (auto s : shape) n *= (size_t)s; return n; }
};

class Conditioner {
public:
    // `get` returns the checkpoint tensor of that name or throws
    Conditioner(const xtts_config& cfg, const std::function<HostTensorView(const std::string&)>& get, cudaStream_t st);
    ~Conditioner();
    // writes cond [n_cond*H] and g [spk_proj] (device pointers)
    void run(const float* wav22k_host, int64_t n22, const float* wav16k_host, int64_t n16, int cond_len_s,
             int chunk_len_s, float* cond_dev, float* g_dev);
    // stage taps for parity tests (host copies of the last run)
    std::vector<float> last_mel;       // [frames,80] of the first piece
    int last_mel_frames = 0;

private:
    struct Impl;
    std::unique_ptr<Impl> impl;
};

}  // namespace xtts
