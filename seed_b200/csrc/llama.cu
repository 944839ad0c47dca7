// llama.cu -- seedb200_llama: models/llama_xformer.py LlamaForCausalLM.forward (:661-743) /
// LlamaModel.forward (:496-627) / LlamaDecoderLayer.forward (:280-332) as a fixed launch sequence.
//
// What changes relative to the reference graph:
//   * q/k/v projections run as ONE GEMM over a fused [3h, h] weight, gate/up as ONE GEMM whose epilogue
//     applies SiLU(gate)*up (weights interleaved in 128-row blocks at create time);
//   * the dense additive mask (:50-92, :552-557) and the per-layer `attention_mask.sum() == 0` host sync
//     (:255) are gone: causality is a kernel flag;
//   * the KV cache is preallocated [B, H, max_seq, D] and appended in place by the RoPE kernel instead of
//     torch.cat per layer per step (:236-237);
//   * batch-1 decode (S == 1) switches to the HBM-bound GEMV / split-KV kernels.
// Weights named like the HF checkpoint (model.layers.N.self_attn.q_proj.weight ...).  o_proj, down_proj,
// norms, embed_tokens and lm_head are borrowed; q/k/v and gate/up are copied into their fused layouts and
// the originals are not referenced after create.
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "ops.h"

namespace sb {
struct LlamaLayerW {
  const __half *in_ln, *post_ln, *qkv_w, *o_w, *gu_w, *down_w;
  __half *k_cache, *v_cache;
};
}  // namespace sb

struct seedb200_llama {
  seedb200_llama_config cfg;
  std::map<std::string, seedb200_tensor> w;
  std::vector<void*> owned;
  const __half *embed, *norm_w, *lm_head;
  std::vector<sb::LlamaLayerW> layers;
  const void *cos_t, *sin_t;
  int max_pos;
  __half *x, *nb, *qkv, *q, *att, *gu, *hn, *last;
  float* da_ws;
  int* da_tickets;
  int last_T;
  int device;
  // ---- device-resident generation loop (seedb200_llama_generate) ----
  int vpad;                       // logits row stride: vocab rounded up to a multiple of 8
  int* gstate;                    // {cache length, step, arrive counter, valid steps, any-unfinished flag}
  int* gfinished;                 // [max_batch]
  int64_t* gtok;                  // [max_batch] token fed to the next decode forward
  int64_t* gout;                  // [max_batch, max_seq] generated tokens
  __half* glogits;                // [max_batch, vpad]
  sb::GenParams* gparams;         // sampling parameters + eos/pad, read by the sampler at run time
  int* gstate_host;               // pinned mirror of gstate (early-stop polling)
  cudaStream_t gstream;           // private stream the decode step is captured on (the caller's may be the legacy
                                  // default stream, which cannot be captured); replays go to the caller's stream
  cudaGraphExec_t gexec[5];       // decode-step graph per batch size (1..4)
  int gunit_launches[5];          // kernels inside one captured unit (launch accounting of graph replays)
  int used_graph;
  int gen_cache_len;
};

namespace sb {

static int llama_find(const seedb200_llama* m, const std::string& name, const __half** out, int64_t n_expected) {
  auto it = m->w.find(name);
  if (it == m->w.end()) {
    set_error("llama_create: missing weight '%s'", name.c_str());
    return SEEDB200_ERR_INVALID;
  }
  const seedb200_tensor& t = it->second;
  int64_t n = 1;
  for (int i = 0; i < t.ndim; ++i) n *= t.shape[i];
  if (t.dtype != SEEDB200_F16 || n != n_expected || (reinterpret_cast<uintptr_t>(t.data) & 15) != 0) {
    set_error("llama_create: weight '%s' must be fp16, 16-byte aligned, %lld elements (got %lld)", name.c_str(),
              (long long)n_expected, (long long)n);
    return SEEDB200_ERR_INVALID;
  }
  *out = static_cast<const __half*>(t.data);
  return 0;
}

template <typename T>
static int llama_alloc(seedb200_llama* m, T** p, size_t elems) {
  void* q = nullptr;
  size_t bytes = elems * sizeof(T);
  SB_CHECK_CUDA(cudaMalloc(&q, bytes < 256 ? 256 : bytes));
  m->owned.push_back(q);
  *p = static_cast<T*>(q);
  return 0;
}

static int llama_build(seedb200_llama* m) {
  const seedb200_llama_config& c = m->cfg;
  const int64_t h = c.hidden, ffn = c.ffn, V = c.vocab;
  cudaStream_t st = 0;
  char nm[256];
  SB_PROPAGATE(llama_find(m, "model.embed_tokens.weight", &m->embed, V * h));
  SB_PROPAGATE(llama_find(m, "model.norm.weight", &m->norm_w, h));
  SB_PROPAGATE(llama_find(m, "lm_head.weight", &m->lm_head, V * h));
  m->layers.resize(c.layers);
  const size_t cache_elems = (size_t)c.max_batch * c.heads * c.max_seq * c.head_dim;
  for (int l = 0; l < c.layers; ++l) {
    LlamaLayerW& L = m->layers[l];
    auto key = [&](const char* s) { snprintf(nm, sizeof(nm), "model.layers.%d.%s", l, s); return std::string(nm); };
    const __half *wq, *wk, *wv, *wg, *wu;
    SB_PROPAGATE(llama_find(m, key("input_layernorm.weight"), &L.in_ln, h));
    SB_PROPAGATE(llama_find(m, key("post_attention_layernorm.weight"), &L.post_ln, h));
    SB_PROPAGATE(llama_find(m, key("self_attn.q_proj.weight"), &wq, h * h));
    SB_PROPAGATE(llama_find(m, key("self_attn.k_proj.weight"), &wk, h * h));
    SB_PROPAGATE(llama_find(m, key("self_attn.v_proj.weight"), &wv, h * h));
    SB_PROPAGATE(llama_find(m, key("self_attn.o_proj.weight"), &L.o_w, h * h));
    SB_PROPAGATE(llama_find(m, key("mlp.gate_proj.weight"), &wg, ffn * h));
    SB_PROPAGATE(llama_find(m, key("mlp.up_proj.weight"), &wu, ffn * h));
    SB_PROPAGATE(llama_find(m, key("mlp.down_proj.weight"), &L.down_w, h * ffn));
    __half *fq, *fg;
    SB_PROPAGATE(llama_alloc(m, &fq, (size_t)3 * h * h));
    SB_CHECK_CUDA(cudaMemcpyAsync(fq, wq, (size_t)h * h * 2, cudaMemcpyDeviceToDevice, st));
    SB_CHECK_CUDA(cudaMemcpyAsync(fq + (size_t)h * h, wk, (size_t)h * h * 2, cudaMemcpyDeviceToDevice, st));
    SB_CHECK_CUDA(cudaMemcpyAsync(fq + (size_t)2 * h * h, wv, (size_t)h * h * 2, cudaMemcpyDeviceToDevice, st));
    L.qkv_w = fq;
    // [ffn/128] blocks of [128 gate rows | 128 up rows]
    SB_PROPAGATE(llama_alloc(m, &fg, (size_t)2 * ffn * h));
    const size_t blk = (size_t)128 * h * 2;
    SB_CHECK_CUDA(cudaMemcpy2DAsync(fg, 2 * blk, wg, blk, blk, ffn / 128, cudaMemcpyDeviceToDevice, st));
    SB_CHECK_CUDA(cudaMemcpy2DAsync(reinterpret_cast<uint8_t*>(fg) + blk, 2 * blk, wu, blk, blk, ffn / 128,
                                    cudaMemcpyDeviceToDevice, st));
    L.gu_w = fg;
    SB_PROPAGATE(llama_alloc(m, &L.k_cache, cache_elems));
    SB_PROPAGATE(llama_alloc(m, &L.v_cache, cache_elems));
  }
  SB_PROPAGATE(get_rope_tables(c.head_dim, c.rope_base, c.max_seq, &m->cos_t, &m->sin_t, &m->max_pos, st));
  const size_t T = (size_t)c.max_batch * c.max_seq;
  SB_PROPAGATE(llama_alloc(m, &m->x, T * h));
  SB_PROPAGATE(llama_alloc(m, &m->nb, T * h));
  SB_PROPAGATE(llama_alloc(m, &m->qkv, T * 3 * h));
  SB_PROPAGATE(llama_alloc(m, &m->q, T * h));
  SB_PROPAGATE(llama_alloc(m, &m->att, T * h));
  SB_PROPAGATE(llama_alloc(m, &m->gu, T * ffn));
  SB_PROPAGATE(llama_alloc(m, &m->hn, T * h));
  SB_PROPAGATE(llama_alloc(m, &m->last, (size_t)c.max_batch * h));
  SB_PROPAGATE(llama_alloc(m, &m->da_ws, (size_t)c.max_batch * c.heads * decode_attention_max_splits(c.max_seq) * (128 + 2)));
  SB_PROPAGATE(llama_alloc(m, &m->da_tickets, (size_t)c.max_batch * c.heads));   // zero now, left zero by every launch
  SB_CHECK_CUDA(cudaMemsetAsync(m->da_tickets, 0, (size_t)c.max_batch * c.heads * sizeof(int), st));
  m->vpad = (c.vocab + 7) / 8 * 8;
  SB_PROPAGATE(llama_alloc(m, &m->gstate, 8));
  SB_PROPAGATE(llama_alloc(m, &m->gfinished, (size_t)c.max_batch));
  SB_PROPAGATE(llama_alloc(m, &m->gtok, (size_t)c.max_batch));
  SB_PROPAGATE(llama_alloc(m, &m->gout, (size_t)c.max_batch * c.max_seq));
  SB_PROPAGATE(llama_alloc(m, &m->glogits, (size_t)c.max_batch * m->vpad));
  SB_PROPAGATE(llama_alloc(m, &m->gparams, 1));
  SB_CHECK_CUDA(cudaMallocHost(reinterpret_cast<void**>(&m->gstate_host), 8 * sizeof(int)));
  SB_CHECK_CUDA(cudaStreamCreateWithFlags(&m->gstream, cudaStreamNonBlocking));
  SB_CHECK_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int get_option(const char* key);

static int lin(cudaStream_t st, int ctas, int M, int N, int K, const void* A, const void* W, void* out, int64_t ldo,
               const void* residual, int mode, const void* norm_w = nullptr, float eps = 0.0f) {
  // norm_w: only on the M <= 4 (decode) path, where the GEMV normalises its activations while staging them
  if (M <= 4) return gemv(A, W, K, out, residual, norm_w, eps, M, N, K, mode, st, ldo);
  seedb200_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = M; d.N = N; d.K = K; d.A = A; d.lda = K; d.W = W; d.ldw = K;
  d.out = out; d.ldo = ldo; d.residual = residual; d.ldr = ldo; d.mode = mode; d.ctas = ctas;
  return gemm(d, st);
}

// dyn (device, optional): {cache length, ...} read by the RoPE/append and decode-attention kernels instead of the
// host value `past_len` -- what makes a captured decode step position independent (S must be 1).
static int llama_forward(seedb200_llama* m, const int64_t* input_ids, const void* inputs_embeds,
                         const int64_t* position_ids, int B, int S, int past_len, int logits_mode, void* logits,
                         int64_t logits_ld, cudaStream_t st, const int* dyn = nullptr) {
  const seedb200_llama_config& c = m->cfg;
  const int h = c.hidden, H = c.heads, D = c.head_dim, ffn = c.ffn, V = c.vocab, ct = c.gemm_ctas;
  const int T = B * S;
  if (input_ids)
    SB_PROPAGATE(embedding(m->embed, h, input_ids, T, h, m->x, h, V, st));
  else
    SB_CHECK_CUDA(cudaMemcpyAsync(m->x, inputs_embeds, (size_t)T * h * 2, cudaMemcpyDeviceToDevice, st));
  const float scale = 1.0f / sqrtf((float)D);   // xformers default scale
  const int kv_len = past_len + S;
  const bool fuse_norm = T <= 4;     // decode: RMSNorm folded into the GEMV that consumes it
  PdlScope pdl(T <= 4 && get_option("decode_pdl") != 0);   // decode chain: programmatic dependent launches
  const bool fused_attn = S == 1 && get_option("decode_fused_attention") != 0 && decode_attention_rope_supported(D, c.max_seq);
  for (int l = 0; l < c.layers; ++l) {
    const LlamaLayerW& L = m->layers[l];
    if (fuse_norm) {
      SB_PROPAGATE(lin(st, ct, T, 3 * h, h, m->x, L.qkv_w, m->qkv, 3 * h, nullptr, 0, L.in_ln, c.rms_eps));
    } else {
      SB_PROPAGATE(rmsnorm(m->x, h, L.in_ln, m->nb, h, T, h, c.rms_eps, st));
      SB_PROPAGATE(lin(st, ct, T, 3 * h, h, m->nb, L.qkv_w, m->qkv, 3 * h, nullptr, 0));
    }
    // qkv rows are [q | k | v] per token, each [H, D]
    if (fused_attn) {
      // cached decode step: RoPE + append + attention in one launch (bit-identical to the pair below up to 512 keys)
      SB_PROPAGATE(decode_attention_rope(m->qkv, position_ids, B, H, D, past_len, c.max_seq, m->max_pos, m->cos_t,
                                         m->sin_t, L.k_cache, L.v_cache, m->att, scale, st, dyn));
    } else {
    SB_PROPAGATE(rope_kv_append_tables(m->qkv, position_ids, B, S, H, D, past_len, c.max_seq, m->max_pos, m->cos_t,
                                       m->sin_t, m->q, L.k_cache, L.v_cache, st, dyn));
    if (S == 1) {
      SB_PROPAGATE(decode_attention(m->q, L.k_cache, L.v_cache, m->att, B, H, D, kv_len, c.max_seq, scale, m->da_ws, st, dyn,
                                    m->da_tickets));
    } else {
      seedb200_attn_desc a;
      memset(&a, 0, sizeof(a));
      a.q = m->q; a.k = L.k_cache; a.v = L.v_cache; a.o = m->att;
      a.q_bs = (int64_t)S * h; a.q_hs = D; a.q_ts = h;
      a.k_bs = (int64_t)H * c.max_seq * D; a.k_hs = (int64_t)c.max_seq * D; a.k_ts = D;
      a.v_bs = a.k_bs; a.v_hs = a.k_hs; a.v_ts = a.k_ts;
      a.o_bs = (int64_t)S * h; a.o_hs = D; a.o_ts = h;
      a.batch = B; a.heads = H; a.nq = S; a.nk = kv_len; a.head_dim = D; a.causal = 1; a.scale = scale;
      SB_PROPAGATE(attention(a, st));
    }
    }
    SB_PROPAGATE(lin(st, ct, T, h, h, m->att, L.o_w, m->x, h, m->x, 0));
    if (fuse_norm) {
      SB_PROPAGATE(lin(st, ct, T, 2 * ffn, h, m->x, L.gu_w, m->gu, ffn, nullptr, 1, L.post_ln, c.rms_eps));
    } else {
      SB_PROPAGATE(rmsnorm(m->x, h, L.post_ln, m->nb, h, T, h, c.rms_eps, st));
      SB_PROPAGATE(lin(st, ct, T, 2 * ffn, h, m->nb, L.gu_w, m->gu, ffn, nullptr, 1));
    }
    SB_PROPAGATE(lin(st, ct, T, h, ffn, m->gu, L.down_w, m->x, h, m->x, 0));
  }
  SB_PROPAGATE(rmsnorm(m->x, h, m->norm_w, m->hn, h, T, h, c.rms_eps, st));
  m->last_T = T;
  if (logits == nullptr) return 0;
  if (logits_mode == 0) {
    SB_PROPAGATE(lin(st, ct, T, V, h, m->hn, m->lm_head, logits, logits_ld, nullptr, 0));
  } else {
    const __half* src = m->hn;
    if (S > 1) {   // gather the last position of every sequence
      SB_CHECK_CUDA(cudaMemcpy2DAsync(m->last, (size_t)h * 2, m->hn + (size_t)(S - 1) * h, (size_t)S * h * 2,
                                      (size_t)h * 2, B, cudaMemcpyDeviceToDevice, st));
      src = m->last;
    }
    SB_PROPAGATE(lin(st, ct, B, V, h, src, m->lm_head, logits, logits_ld, nullptr, 0));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// generation loop on the device (scripts/seed_llama_inference_8B.py:26-38 -> HF sample/greedy_search +
// llama_xformer.py:745-776): prefill, then max_new_tokens x (sampler -> cached decode forward).
// ---------------------------------------------------------------------------------------------------------------
__global__ void gen_reset_kernel(GenParams gp, GenParams* dst, int* state, int* finished, int past_len, int B) {
  if (threadIdx.x == 0) {
    *dst = gp;
    state[0] = past_len; state[1] = 0; state[2] = 0; state[3] = 0; state[4] = 0;
  }
  if (threadIdx.x < B) finished[threadIdx.x] = 0;
}

// one unit of the loop: cached forward of the tokens in m->gtok, then the sampler (which also advances the device
// counters).  Every launch argument is a handle-owned pointer or a constant: the unit can be captured once.
static int gen_unit(seedb200_llama* m, int B, cudaStream_t st) {
  PdlScope pdl(get_option("decode_pdl") != 0);
  SB_PROPAGATE(llama_forward(m, m->gtok, nullptr, nullptr, B, 1, 0, 1, m->glogits, m->vpad, st, m->gstate));
  SB_PROPAGATE(sample(m->glogits, m->vpad, B, m->cfg.vocab, nullptr, m->gparams, 0, m->gstate, /*advance_cache=*/1,
                      m->gtok, m->gout, m->cfg.max_seq, m->gfinished, st));
  return 0;
}

static int gen_capture(seedb200_llama* m, int B) {
  cudaGraph_t graph = nullptr;
  cudaStream_t st = m->gstream;
  SB_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int64_t before = seedb200_launch_count();
  const int s = gen_unit(m, B, st);
  const int unit = (int)(seedb200_launch_count() - before);
  count_launch(-unit);                        // captured, not executed
  const cudaError_t e = cudaStreamEndCapture(st, &graph);
  if (s != 0) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return s;
  }
  if (e != cudaSuccess || graph == nullptr) {
    set_error("llama_generate: stream capture of the decode step failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return SEEDB200_ERR_CUDA;
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ei != cudaSuccess) {
    set_error("llama_generate: cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
    cudaGetLastError();
    return SEEDB200_ERR_CUDA;
  }
  m->gexec[B] = exec;
  m->gunit_launches[B] = unit;
  return 0;
}

static int llama_generate(seedb200_llama* m, const int64_t* prompt_ids, int B, int S, int max_new,
                          const seedb200_sample_params* sp, int64_t eos, int64_t pad, int use_graph,
                          int64_t* tokens_out, int* n_generated_host, cudaStream_t st) {
  const seedb200_llama_config& c = m->cfg;
  GenParams gp;
  gp.sp = *sp; gp.eos = eos; gp.pad = pad;
  gen_reset_kernel<<<1, 32, 0, st>>>(gp, m->gparams, m->gstate, m->gfinished, S, B);
  SB_LAUNCH_CHECK();
  // prefill: cache rows [0, S), logits of the last position only (what the sampler needs)
  SB_PROPAGATE(llama_forward(m, prompt_ids, nullptr, nullptr, B, S, 0, 1, m->glogits, m->vpad, st));
  // token 0 comes from the prefill logits: the step counter advances, the cache length does not
  SB_PROPAGATE(sample(m->glogits, m->vpad, B, c.vocab, nullptr, m->gparams, 0, m->gstate, /*advance_cache=*/0, m->gtok,
                      m->gout, c.max_seq, m->gfinished, st));
  int units = max_new - 1, done = 0;
  m->used_graph = 0;
  auto all_finished = [&](bool* stop) -> int {   // early stop: poll the device counters (only with an eos id)
    SB_CHECK_CUDA(cudaMemcpyAsync(m->gstate_host, m->gstate, 5 * sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CHECK_CUDA(cudaStreamSynchronize(st));
    *stop = m->gstate_host[3] < m->gstate_host[1];      // a step ran with every sequence already finished
    return 0;
  };
  if (units > 0) {   // first unit eagerly (also resolves every lazily-set function attribute before a capture)
    SB_PROPAGATE(gen_unit(m, B, st));
    done = 1;
  }
  if (units > done && use_graph) {
    if (m->gexec[B] == nullptr) {
      int s = gen_capture(m, B);
      if (s != 0 && get_option("decode_pdl") != 0) {   // retry without programmatic launches inside the graph
        seedb200_set_option("decode_pdl", 0);
        s = gen_capture(m, B);
        seedb200_set_option("decode_pdl", 1);
      }
      if (s != 0) return s;
    }
    m->used_graph = 1;
  }
  bool stop = false;
  while (done < units && !stop) {
    int chunk = units - done;
    if (eos >= 0 && chunk > 32) chunk = 32;
    for (int i = 0; i < chunk; ++i) {
      if (m->used_graph) {
        SB_CHECK_CUDA(cudaGraphLaunch(m->gexec[B], st));
        count_launch(m->gunit_launches[B]);
      } else {
        SB_PROPAGATE(gen_unit(m, B, st));
      }
    }
    done += chunk;
    if (eos >= 0 && done < units) SB_PROPAGATE(all_finished(&stop));
  }
  int n_valid = done + 1;
  if (eos >= 0) {
    SB_PROPAGATE(all_finished(&stop));
    n_valid = m->gstate_host[3];
  }
  if (tokens_out != nullptr)
    SB_CHECK_CUDA(cudaMemcpy2DAsync(tokens_out, (size_t)max_new * 8, m->gout, (size_t)c.max_seq * 8, (size_t)n_valid * 8, B,
                                    cudaMemcpyDeviceToDevice, st));
  if (n_generated_host != nullptr) *n_generated_host = n_valid;
  m->gen_cache_len = S + done;     // tokens whose K/V are in the cache
  return 0;
}

}  // namespace sb

extern "C" {

int seedb200_llama_create(const seedb200_llama_config* cfg, const seedb200_tensor* weights, int n_weights,
                          seedb200_llama** out) {
  if (!cfg || !weights || !out) {
    sb::set_error("llama_create: null argument");
    return SEEDB200_ERR_INVALID;
  }
  SB_REQUIRE(cfg->head_dim == 128, "llama_create: head_dim must be 128 (got %d)", cfg->head_dim);
  SB_REQUIRE(cfg->hidden == cfg->heads * cfg->head_dim, "llama_create: hidden != heads * head_dim");
  SB_REQUIRE(cfg->ffn % 128 == 0, "llama_create: ffn %d must be a multiple of 128", cfg->ffn);
  SB_REQUIRE(cfg->hidden % 8 == 0 && cfg->layers >= 0 && cfg->vocab > 0, "llama_create: bad dims");
  SB_REQUIRE(cfg->max_batch >= 1 && cfg->max_seq >= 1, "llama_create: bad cache size");
  seedb200_llama* m = new seedb200_llama();
  m->cfg = *cfg;
  if (m->cfg.rope_base <= 0.0f) m->cfg.rope_base = 10000.0f;
  if (m->cfg.rms_eps <= 0.0f) m->cfg.rms_eps = 1e-6f;
  m->last_T = 0;
  m->device = sb::cur_device();
  m->gstate_host = nullptr;
  m->gstream = nullptr;
  m->used_graph = -1;
  m->gen_cache_len = 0;
  for (int i = 0; i < 5; ++i) { m->gexec[i] = nullptr; m->gunit_launches[i] = 0; }
  for (int i = 0; i < n_weights; ++i) m->w[std::string(weights[i].name)] = weights[i];
  int s = sb::llama_build(m);
  if (s != 0) {
    seedb200_llama_destroy(m);
    return s;
  }
  m->w.clear();
  *out = m;
  return 0;
}

void seedb200_llama_destroy(seedb200_llama* llm) {
  if (!llm) return;
  for (int i = 0; i < 5; ++i)
    if (llm->gexec[i]) cudaGraphExecDestroy(llm->gexec[i]);
  if (llm->gstate_host) cudaFreeHost(llm->gstate_host);
  if (llm->gstream) cudaStreamDestroy(llm->gstream);
  for (void* p : llm->owned) cudaFree(p);
  delete llm;
}

int seedb200_llama_forward_ld(seedb200_llama* llm, const int64_t* input_ids, const void* inputs_embeds,
                              const int64_t* position_ids, int B, int S, int past_len, int logits_mode,
                              void* logits_out, int64_t logits_ld, void* stream) {
  SB_REQUIRE(llm != nullptr, "llama_forward: null handle");
  SB_REQUIRE((input_ids != nullptr) != (inputs_embeds != nullptr),
             "llama_forward: specify exactly one of input_ids / inputs_embeds (llama_xformer.py:516-523)");
  SB_REQUIRE(B >= 1 && B <= llm->cfg.max_batch, "llama_forward: batch %d outside [1,%d]", B, llm->cfg.max_batch);
  SB_REQUIRE(S >= 1 && past_len >= 0 && past_len + S <= llm->cfg.max_seq,
             "llama_forward: past_len %d + S %d exceeds max_seq %d", past_len, S, llm->cfg.max_seq);
  SB_REQUIRE(logits_mode == 0 || logits_mode == 1, "llama_forward: logits_mode must be 0 or 1");
  SB_REQUIRE(logits_out == nullptr || logits_ld >= llm->cfg.vocab, "llama_forward: logits_ld %lld < vocab %d",
             (long long)logits_ld, llm->cfg.vocab);
  sb::DeviceGuard guard(llm->device);
  return sb::llama_forward(llm, input_ids, inputs_embeds, position_ids, B, S, past_len, logits_mode, logits_out,
                           logits_ld, static_cast<cudaStream_t>(stream));
}

int seedb200_llama_forward(seedb200_llama* llm, const int64_t* input_ids, const void* inputs_embeds,
                           const int64_t* position_ids, int B, int S, int past_len, int logits_mode, void* logits_out,
                           void* stream) {
  return seedb200_llama_forward_ld(llm, input_ids, inputs_embeds, position_ids, B, S, past_len, logits_mode, logits_out,
                                   llm ? llm->cfg.vocab : 0, stream);
}

int seedb200_llama_generate(seedb200_llama* llm, const int64_t* prompt_ids, int B, int S, int max_new_tokens,
                            const seedb200_sample_params* sp, int64_t eos_id, int64_t pad_id, int use_graph,
                            int64_t* tokens_out, int* n_generated_host, void* stream) {
  SB_REQUIRE(llm && prompt_ids && sp, "llama_generate: null argument");
  SB_REQUIRE(B >= 1 && B <= llm->cfg.max_batch && B <= 4, "llama_generate: batch %d outside [1,%d] (decode GEMV: <= 4 rows)",
             B, llm->cfg.max_batch < 4 ? llm->cfg.max_batch : 4);
  SB_REQUIRE(S >= 1 && max_new_tokens >= 1 && S + max_new_tokens <= llm->cfg.max_seq,
             "llama_generate: prompt %d + %d new tokens exceeds max_seq %d", S, max_new_tokens, llm->cfg.max_seq);
  SB_REQUIRE(!sp->do_sample || (sp->temperature > 0.0f && sp->top_p > 0.0f), "llama_generate: temperature and top_p must be > 0");
  sb::DeviceGuard guard(llm->device);
  return sb::llama_generate(llm, prompt_ids, B, S, max_new_tokens, sp, eos_id, pad_id, use_graph, tokens_out,
                            n_generated_host, static_cast<cudaStream_t>(stream));
}

int seedb200_llama_generate_used_graph(seedb200_llama* llm) { return llm ? llm->used_graph : -1; }

int seedb200_llama_kv_ptrs(seedb200_llama* llm, int layer, void** k, void** v) {
  SB_REQUIRE(llm && k && v && layer >= 0 && layer < (int)llm->layers.size(), "llama_kv_ptrs: bad arguments");
  *k = llm->layers[layer].k_cache;
  *v = llm->layers[layer].v_cache;
  return 0;
}

int seedb200_llama_kv_load(seedb200_llama* llm, int layer, const void* k, const void* v, int B, int past_len,
                           void* stream) {
  SB_REQUIRE(llm && k && v && layer >= 0 && layer < (int)llm->layers.size(), "llama_kv_load: bad arguments");
  SB_REQUIRE(B >= 1 && B <= llm->cfg.max_batch && past_len >= 0 && past_len <= llm->cfg.max_seq, "llama_kv_load: bad sizes");
  if (past_len == 0) return 0;
  sb::DeviceGuard guard(llm->device);
  const size_t D = llm->cfg.head_dim, H = llm->cfg.heads;
  const size_t w = (size_t)past_len * D * 2, dp = (size_t)llm->cfg.max_seq * D * 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // source [B,H,past,D] contiguous; destination rows of max_seq*D per (b,h); batch stride H*max_seq*D matches
  SB_CHECK_CUDA(cudaMemcpy2DAsync(llm->layers[layer].k_cache, dp, k, w, w, (size_t)B * H, cudaMemcpyDeviceToDevice, st));
  SB_CHECK_CUDA(cudaMemcpy2DAsync(llm->layers[layer].v_cache, dp, v, w, w, (size_t)B * H, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int64_t seedb200_llama_tap(seedb200_llama* llm, int what, void* dst, int64_t max_elems, void* stream) {
  if (!llm || !dst || llm->last_T <= 0 || what != 0) return -1;
  sb::DeviceGuard guard(llm->device);
  int64_t n = (int64_t)llm->last_T * llm->cfg.hidden;
  if (n > max_elems) n = max_elems;
  if (cudaMemcpyAsync(dst, llm->hn, (size_t)n * 2, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)) != cudaSuccess)
    return -1;
  return n;
}

}  // extern "C"
