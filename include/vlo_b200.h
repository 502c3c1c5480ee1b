/*
 * vlo_b200.h — C ABI of libvlo_b200.so, the B200-native (sm_100a) engine for the per-frame hot
 * loop of showlab/videollm-online.
 *
 * The reference has no FFI: its seam is the Python duck-typed API of demo/inference.py and
 * models/*.py.  Each entry point below names the reference call it replaces (path:line relative
 * to the reference checkout, `HF:` = transformers 5.5).  INTEGRATION.md shows the ctypes binding
 * a maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns 0 on success, a negative value on error; vlo_last_error() returns a
 *    thread-local message.  No exceptions cross the boundary.
 *  - pointers named d_* are raw DEVICE pointers (torch `tensor.data_ptr()`); h_* are host pointers.
 *    The engine never frees caller memory; the caller never frees engine memory.
 *  - `cuda_stream` is a cudaStream_t passed as void*; calls are asynchronous w.r.t. the host unless
 *    documented otherwise.
 *  - one engine per GPU/process; calls on one engine are not thread-safe (the Python host holds a lock).
 *  - 16-bit tensors: decoder/connector = bf16, vision tower = fp16 (the reference runs the ViT under
 *    fp16 autocast: models/modeling_live.py:23, models/vision_live.py:13).
 */
#ifndef VLO_B200_H_
#define VLO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vlo_engine vlo_engine;

/* Model + capacity description.  Field meaning follows LlamaConfig / SiglipVisionConfig /
 * LiveConfigMixin (models/configuration_live.py:5-21). */
typedef struct vlo_config {
  /* decoder (Llama) */
  int32_t hidden_size, num_layers, num_heads, num_kv_heads, head_dim, intermediate_size, vocab_size;
  float rms_norm_eps;
  /* vision tower (SigLIP) */
  int32_t vit_hidden, vit_layers, vit_heads, vit_mlp, image_size, patch_size;
  float vit_ln_eps;
  /* frame tokens: frame_token_cls, frame_token_pooled (models/arguments_live.py:41-47) */
  int32_t frame_token_cls, pool_h, pool_w;
  /* capacities */
  int32_t max_streams;     /* concurrent video streams (KV caches) on this GPU */
  int32_t max_kv_tokens;   /* per-stream KV capacity, tokens */
  int32_t max_step_tokens; /* max total new tokens in one vlo_step */
  int32_t max_vit_batch;   /* max frames per vlo_vit_encode */
} vlo_config;

/* Per-sequence result of a decoder step: everything demo/inference.py:76-81 and
 * models/modeling_live.py:177-179 read back from the logits, computed on the device. */
typedef struct vlo_decision {
  int32_t argmax_id;          /* plain greedy id (fast_greedy_generate) */
  int32_t argmax_excl_id;     /* argmax with the interval id removed (p(interval) < threshold case) */
  float p_interval;           /* softmax probability of the interval id, rounded to bf16 like the reference */
  float max_logit;            /* for diagnostics / margins */
  float top2_margin;          /* max logit - second max logit */
  float lse;                  /* log-sum-exp of the logits */
  int32_t argmax_prob_id;     /* argmax over the bf16-rounded softmax (what next_score.argmax sees), first index wins ties */
  int32_t reserved1;
} vlo_decision;

const char* vlo_last_error(void);
/* number of kernels of this library launched so far in this process (bench: gpu_launches) */
long long vlo_launch_count(void);
/* Optional CUDA-event profiler: when enabled every launch of the classes below is bracketed by an event
 * pair on its stream.  vlo_profile_read synchronises, returns per-class sums since the last read
 * (milliseconds, launches, algorithmic HBM bytes as defined in DESIGN.md) and resets.
 * classes: 0 weight-streaming GEMM (decoder/connector/lm_head), 1 KV-append attention, 2 split-KV merge,
 *          3 ViT GEMM, 4 ViT attention, 5 other. */
#define VLO_PROF_NUM_CLASSES 6
int vlo_profile_enable(int on);
int vlo_profile_read(double* ms, long long* launches, double* algo_bytes, int n_classes);
/* 1 when the device is an sm_100 part the kernels can run on */
int vlo_device_supported(int device);

/* ------------------------------------------------------------------ engine life cycle
 * replaces: build_live / LiveLlamaForCausalLM.from_pretrained + model.to('cuda')
 *           (models/modeling_live.py:184-222, demo/inference.py:15-16) */
int vlo_engine_create(const vlo_config* cfg, int device, vlo_engine** out);
int vlo_engine_destroy(vlo_engine* e);
/* Register one weight tensor that already lives on the device in the ENGINE layout (see
 * DESIGN.md "weight layout"; the Python loader does the HF/peft key mapping, LoRA merge and
 * packing at init).  The engine keeps the pointer; the caller keeps the storage alive. */
int vlo_load_tensor(vlo_engine* e, const char* name, const void* d_ptr, int64_t n_bytes);
int vlo_finalize_weights(vlo_engine* e);
/* bytes of device memory the engine allocated itself (KV + workspaces) */
int64_t vlo_engine_device_bytes(vlo_engine* e);

/* ------------------------------------------------------------------ streams (KV ownership)
 * replaces: the DynamicCache handed around as past_key_values (HF:cache_utils.py:88-121;
 *           demo/inference.py:47,69-70,91) */
int vlo_stream_open(vlo_engine* e, int* stream_id);
int vlo_stream_reset(vlo_engine* e, int stream_id);                 /* LiveInfer.reset: past_key_values = None */
int vlo_stream_close(vlo_engine* e, int stream_id);
int vlo_kv_len(vlo_engine* e, int stream_id, int* out_len);          /* Cache.get_seq_length() */
int vlo_kv_truncate(vlo_engine* e, int stream_id, int new_len);      /* trim_past_key_values(0, new_len), models/modeling_live.py:170-171 */
/* dst stream := the first n_tokens cache positions of src (device copy, all layers); dst's previous contents are dropped.
 * The engine-side form of `trim_past_key_values(past_key_values, 0, n)` when the ORIGINAL cache must survive
 * (stream_evaluate's look-ahead, models/modeling_live.py:112-113,170-171). */
int vlo_kv_copy_prefix(vlo_engine* e, int src_stream_id, int dst_stream_id, int n_tokens, void* cuda_stream);
/* test/bench hook: fill a stream's KV cache with deterministic pseudo-random values up to n_tokens
 * (pre-fill for the 12k-context measurements without replaying 1200 frames) */
int vlo_kv_fill_synthetic(vlo_engine* e, int stream_id, int n_tokens, uint64_t seed, void* cuda_stream);
/* copy one layer's K or V rows [n_kv_heads, len, head_dim] bf16 out of / into the cache (tests) */
int vlo_kv_read(vlo_engine* e, int stream_id, int layer, int is_v, void* d_out, void* cuda_stream);
int vlo_kv_write(vlo_engine* e, int stream_id, int layer, int is_v, const void* d_in, int n_tokens, void* cuda_stream);

/* ------------------------------------------------------------------ hot path
 * vlo_vit_encode replaces LiveMixin.visual_embed (models/modeling_live.py:21-27) =
 *   _siglip_vision_encode (models/vision_live.py:10-30) + connector
 *   (models/live_llama/modeling_live_llama.py:18-22).
 * d_frames: uint8 [n_frames, 3, image, image]; d_out: bf16 [n_frames * frame_num_tokens, hidden].
 * d_vit_tokens (optional, may be NULL): fp32 [n_frames, frame_num_tokens, vit_hidden] = the
 * pre-connector tokens (what build_live_vision's encode_fn returns). */
int vlo_vit_encode(vlo_engine* e, const uint8_t* d_frames, int n_frames, void* d_out, float* d_vit_tokens,
                   void* cuda_stream);
/* connector only: visual_embed without a vision tower (models/modeling_live.py:22-27 with
 * pre-extracted features). d_tokens: bf16 [n_rows, vit_hidden] */
int vlo_connector(vlo_engine* e, const void* d_tokens, int n_rows, void* d_out, void* cuda_stream);
/* replaces model.get_input_embeddings()(ids) (demo/inference.py:46,66; models/modeling_live.py:181).
 * ids are clamped to vocab-1 like joint_embed does (models/modeling_live.py:38). */
int vlo_embed_tokens(vlo_engine* e, const int64_t* d_ids, int n, void* d_out, void* cuda_stream);
/* KV-append forward for a ragged batch — replaces LiveLlamaForCausalLM.forward(inputs_embeds=...,
 * past_key_values=..., use_cache=True) (models/live_llama/modeling_live_llama.py:24-67 ->
 * HF:models/llama/modeling_llama.py:375-499) plus the read of logits[:, -1:]
 * (demo/inference.py:76-81, models/modeling_live.py:177).
 *   h_stream_ids[n_seqs], h_q_lens[n_seqs]: host arrays; d_embeds: bf16 [sum(q_lens), hidden] packed.
 *   d_last_logits: bf16 [n_seqs, vocab] or NULL; d_decisions: vlo_decision[n_seqs] (device) or NULL.
 * Appends q_lens[i] tokens to stream i's KV cache. */
int vlo_step(vlo_engine* e, int n_seqs, const int32_t* h_stream_ids, const int32_t* h_q_lens, const void* d_embeds,
             void* d_last_logits, vlo_decision* d_decisions, int interval_id, void* cuda_stream);
/* Same, with token rows gathered on the device: d_row_ids[sum(q_lens)] (device int64), one entry per packed row;
 * id >= 0 = "this row is the embedding of token id" (gathered from the table, clamped to vocab-1),
 * id <  0 = "row already present in d_embeds" (a frame embedding).  The steady-state frame step is
 * [interval id, -1 x 10]; an AR step is [last id]; no separate embedding launch or torch.cat. */
int vlo_step_ids(vlo_engine* e, int n_seqs, const int32_t* h_stream_ids, const int32_t* h_q_lens,
                 const int64_t* d_row_ids, const void* d_embeds, void* d_last_logits, vlo_decision* d_decisions,
                 int interval_id, void* cuda_stream);
/* all-position logits for the last vlo_step's tokens (the reference's logits_to_keep=0 behaviour,
 * HF:models/llama/modeling_llama.py:485-487); d_logits bf16 [sum(q_lens), vocab]. Test/compat path. */
int vlo_last_step_logits(vlo_engine* e, void* d_logits, void* cuda_stream);
/* final hidden states (post model.norm) of the last step, bf16 [sum(q_lens), hidden] (tests) */
int vlo_last_step_hidden(vlo_engine* e, void* d_hidden, void* cuda_stream);

/* ------------------------------------------------------------------ measurement hooks (bench.py roofline)
 * Back-to-back launches of one kernel class on the engine's own buffers: the KV-append attention over every
 * layer's cache of the given streams (n_tok query rows at the end of each cache; skip_merge = 1 leaves out the
 * split-KV merge kernel), and the four weight-streaming GEMMs
 * (qkv, o, gate_up, down) of every layer for n_tok token rows.  Asynchronous; the caller brackets the call with
 * CUDA events.  *h_algo_bytes = algorithmic HBM bytes per attention launch / per full pass over the layers. */
int vlo_bench_attn(vlo_engine* e, int n_seqs, const int32_t* h_stream_ids, int n_tok, int iters, int skip_merge,
                   double* h_algo_bytes_per_launch, void* cuda_stream);
int vlo_bench_gemm(vlo_engine* e, int n_tok, int iters, double* h_algo_bytes_per_iter, int* h_launches_per_iter,
                   void* cuda_stream);

/* ------------------------------------------------------------------ kernel-level entry points
 * (used by the parity tests and micro-benchmarks; same kernels the engine launches) */
/* C = A[rows_a,k] * B[rows_b,k]^T on tcgen05; see csrc/gemm.cuh for fmt/epi/act codes */
int vlo_op_gemm(int fmt, int swap, int epi, int act, const void* d_a, int rows_a, const void* d_b, int rows_b, int k,
                void* d_out, int ld_out, const float* d_bias, const float* d_pos, int pos_rows, int splits,
                long long split_stride, int bn, void* cuda_stream);
/* Persistent weight-streaming GEMM (csrc/gemm_ws.cuh): out[t, n] = sum_k X[t,k] W[n,k].  bn = token-tile width
 * (0: smallest of 16/32/64/128 covering rows_x, which must then be <= 128; 64/96/128/192 tile larger rows_x).
 * mode 0: stream-K over all SMs, fp32 partial planes out[plane][rows_x][rows_w] (the caller zero-fills d_out
 *         and sums the planes; *h_max_planes tells how many planes to provide); d_out == NULL only plans.
 * mode 1: whole tiles, 16-bit out[rows_x][ld_out] = act(acc + bias). */
int vlo_op_gemm_ws(int fmt, int mode, const void* d_w, int rows_w, const void* d_x, int rows_x, int k, void* d_out,
                   int ld_out, long long plane_stride, const float* d_bias, int act, int n_ctas, int bn, int* h_max_planes,
                   void* cuda_stream);
/* 2-CTA (tcgen05 cta_group::2) tensor-bound GEMM of the ViT trunk at batch >= 3 (csrc/gemm2.cuh), fp16:
 *   epi 0: out16[rows_x][ld_out] = act(fp16(X W^T + bias));  epi 1: out32[rows_x][ld_out] += fp16(X W^T + bias)
 *   (the fp32 residual stream).  bn = 256 | 128 features per CTA-pair tile; rows_w % 32 == 0; d_bias required.
 * Replaces the cuBLAS GEMMs of HF:models/siglip/modeling_siglip.py:285-287, 309, 323-327. */
int vlo_op_gemm2(const void* d_x, int rows_x, const void* d_w, int rows_w, int k, void* d_out, int ld_out, const float* d_bias,
                 int act, int epi, int bn, void* cuda_stream);
/* KV-append attention over one layer's cache (the graded kernel, K15):
 *   d_q bf16 [n_tok, n_heads, head_dim] (RoPE applied); d_k/d_v bf16 [n_kv_heads, kv_stride, head_dim];
 *   keys 0..kv_len-1 valid, the n_tok query tokens sit at positions kv_len-n_tok .. kv_len-1 (causal with
 *   offset, HF:masking_utils.py:263-272); d_out bf16 [n_tok, n_heads*head_dim]; d_ws fp32 scratch of
 *   vlo_op_attn_ws_bytes(). Replaces HF:models/llama/modeling_llama.py:272-285 (SDPA / flash-attn 2). */
/* which generation of the KV-append attention kernel runs for this head layout: 3 = tcgen05 with P in TMEM and a
 * 192 KB K/V ring (csrc/attn_tc2.cuh, default), 2 = tcgen05 with P in shared memory (csrc/attn_tc.cuh, VLO_ATTN=2),
 * 1 = mma.sync (csrc/attn.cuh; forced by VLO_ATTN=1, or when n_heads/n_kv_heads does not divide 128) */
int vlo_op_attn_version(int n_heads, int n_kv_heads);
/* developer aid: with VLO_ATTN_TRACE=1 the tcgen05 attention kernel records clock64 stamps per CTA and role
 * ([cta][3 roles][64]); this copies the first n values to the host (tools/gpu_attn_trace.py). */
int vlo_debug_attn_trace(long long* h_out, int n);
int64_t vlo_op_attn_ws_bytes(int n_tok, int n_heads, int head_dim, int kv_len);
int vlo_op_attn_kvappend(const void* d_q, const void* d_k, const void* d_v, void* d_out, float* d_ws, int n_tok,
                         int n_heads, int n_kv_heads, int head_dim, int kv_len, long long kv_stride, void* cuda_stream);
/* Micro-loop of the same attention over n_layers separate K / V matrices (working set > L2), iters passes of back-to-back
 * launches on one plan: what bench tools bracket with ONE CUDA-event pair (per-launch time incl. the PDL overlap of a real
 * step).  skip_merge != 0 launches the main kernel only.  Mirrors vlo_bench_attn without an engine. */
int vlo_op_attn_bench(const void* d_q, const void* d_k, const void* d_v, void* d_out, float* d_ws, int n_tok, int n_heads,
                      int n_kv_heads, int head_dim, int kv_len, long long kv_stride, int n_layers,
                      long long layer_stride_rows, int iters, int skip_merge, double* h_algo_bytes, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* VLO_B200_H_ */
