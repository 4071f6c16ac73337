/* wavenet_b200.h -- C ABI of libwavenet_b200.so (sm_100a).
 *
 * Drop-in boundary for the two hot paths of vincentherrmann/pytorch-wavenet:
 *   (T) the training-time dilated causal convolution stack  WaveNetModel.forward / .wavenet
 *       (reference wavenet_model.py:125-196, wavenet_modules.py:10-39,80-127), and
 *   (G) the Fast-WaveNet cached-queue sampling loop         WaveNetModel.generate_fast
 *       (reference wavenet_model.py:237-315, wavenet_modules.py:42-77).
 * The reference is pure Python on torch and has no FFI of its own; these entry points are what a binding
 * for those Python methods calls (see INTEGRATION.md for the ctypes stub a maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t value, or a negative WN_E_* argument error;
 *     nothing throws; wn_last_error_string() describes the last failure on the calling thread.
 *   - all pointers named d_* are DEVICE pointers owned by the caller; no function allocates persistent device
 *     memory behind the caller's back (the sampler handle owns only host-side bookkeeping).
 *   - every launch is asynchronous on the cudaStream_t passed as `void* stream` (NULL = legacy default stream).
 *   - "frames layout": activations are (B, L, C) fp32, C contiguous, frame t of sequence b at ((b*L+t)*C);
 *     time is ABSOLUTE (frame L-1 is the newest sample); a layer's valid frames are [start, L) and history
 *     left of `start` reads as zero -- this restates the reference's left zero-pad + time->batch fold
 *     (wavenet_modules.py:24-37) without moving data.
 */
#ifndef WAVENET_B200_H
#define WAVENET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WN_ABI_VERSION 2

/* operand precision of the tensor-core training kernels (wn_tb_*): what the MATRIX PRODUCTS see; the residual stream and
 * skip stay fp32-class and accumulation is fp32 in both */
#define WN_PREC_BF16       1   /* single-pass bf16 operands (BASELINE.json configs[4], "bf16 training")             */
#define WN_PREC_BF16_PAIRS 2   /* bf16 (hi, lo) pairs, three MMAs per product: fp32-class, the 1e-4 parity path       */

#define WN_E_BADARG   (-1)   /* null pointer / non-positive size / inconsistent shapes        */
#define WN_E_UNSUPP   (-2)   /* shape outside what the kernels cover (message says which)     */
#define WN_E_NODEVICE (-3)   /* no sm_100 device visible                                      */
#define WN_E_STATE    (-4)   /* sampler handle used before wn_gen_bind / after destroy        */

/* ---------------------------------------------------------------- library / device */
int         wn_version(void);
const char* wn_last_error_string(void);
/* sm_count, compute capability, max opt-in shared memory per block, L2 bytes of the CURRENT device */
int         wn_device_info(int* sm_count, int* cc_major, int* cc_minor, int* smem_optin, int* l2_bytes);

/* ---------------------------------------------------------------- (T) weight packing
 * The training kernels read weights K-outer ("transposed") so a K-slab is a contiguous copy into shared
 * memory.  Packing is device->device, asynchronous, and must be redone after the parameters change.
 *
 *   filter/gate  (D,R,k) x2 + biases -> wfg_t [k*R][N1p],  bfg [N1p],  N1p = wn_n1p(D)
 *        column chunk c (128 wide): cols [0,64) = filter channels 64c.., cols [64,128) = gate channels 64c..
 *        row j*R + r holds tap j (j = 0 is the OLDEST tap, as in nn.Conv1d weight[:,:,0]) of input channel r
 *   residual (R,D,1) + skip (S,D,1) + biases -> wrs_t [D][N2p], brs [N2p], N2p = wn_n2p(R+S)
 *        cols [0,R) residual outputs, [R,R+S) skip outputs
 *   plain 1x1 (N,K,1) + bias -> w_t [K][Np], b [Np], Np = wn_n2p(N)          (end_conv_1 / end_conv_2)
 * Bias pointers may be NULL (bias=False in the reference ctor, wavenet_model.py:39): packed bias is zero. */
int wn_n1p(int D);
int wn_n2p(int N);
int wn_pack_gate_weights(const float* d_wf, const float* d_wg, const float* d_bf, const float* d_bg,
                         int R, int D, int k, float* d_wfg_t, float* d_bfg, void* stream);
int wn_pack_res_skip_weights(const float* d_wr, const float* d_ws, const float* d_br, const float* d_bs,
                             int R, int D, int S, float* d_wrs_t, float* d_brs, void* stream);
int wn_pack_1x1_weights(const float* d_w, const float* d_b, int N, int K, float* d_w_t, float* d_b_p, void* stream);

/* ---------------------------------------------------------------- (T) start conv
 * replaces: start_conv applied to the (B,classes,L) input, wavenet_model.py:65-68,127.
 * dense : d_x (B,classes,L) fp32 (any values; one-hot in the reference data path, audio_data.py:119-121)
 * index : d_idx (B,L) class indices (uint8 / int64) -- equals the dense form on one-hot input bit for bit.
 * d_w_t / d_b_p: start_conv.weight (R,classes,1) and bias packed by wn_pack_1x1_weights (N=R, K=classes), i.e.
 * a (classes, wn_n2p(R)) table whose row c is the embedding of class c.  Output d_h (B,L,R) frames layout. */
int wn_start_fwd_dense(const float* d_x, const float* d_w_t, const float* d_b_p, float* d_h,
                       int B, int classes, int L, int R, void* stream);
int wn_start_fwd_index_u8(const uint8_t* d_idx, const float* d_w_t, const float* d_b_p, float* d_h,
                          int B, int classes, int L, int R, void* stream);
int wn_start_fwd_index_i64(const int64_t* d_idx, const float* d_w_t, const float* d_b_p, float* d_h,
                           int B, int classes, int L, int R, void* stream);

/* ---------------------------------------------------------------- (T) one residual block, one launch
 * replaces the loop body of WaveNetModel.wavenet, wavenet_model.py:142-165, including both dilate() calls
 * (wavenet_modules.py:10-39) and the constant pad (:80-127):
 *     z[t]     = tanh(sum_j Wf[:,:,j] hp[t-(k-1-j)d] + bf) * sigmoid(sum_j Wg[:,:,j] hp[t-(k-1-j)d] + bg)
 *     h_out[t] = Wr z[t] + br + hp[t]                     for t in [out_start, L)
 *     skip[t]  = Ws z[t] + bs (+ skip[t] unless skip_init) for t in [skip_start, L)
 * with hp[t] = h_in[t] for t >= in_start, else 0.  d_skip is (B, L-skip_start, S) frames layout.
 * mode: 0 = exact fp32 FFMA (any shape); other values are reserved for the tensor-core variants. */
typedef struct wn_block_args {
    const float* d_h_in;  float* d_h_out;  float* d_skip;
    const float* d_wfg_t; const float* d_bfg; const float* d_wrs_t; const float* d_brs;
    int B, L, R, D, S, k, dilation;
    int in_start, out_start, skip_start, skip_init;
    int mode;
    float* d_fg_save;   /* optional (B,L,2D): tanh(F) in [0,D), sigmoid(G) in [D,2D) per frame, kept for the backward */
} wn_block_args;
int wn_block_fwd(const wn_block_args* a, void* stream);

/* ---------------------------------------------------------------- (T) the same block on the tensor cores
 * tcgen05.mma kind::tf32 with 3xTF32 operand splitting (hi*hi + lo*hi + hi*lo, fp32 accumulation in tensor memory):
 * fp32-class accuracy (~1e-6 relative) at tensor-core rate.  Two launches per block (conv+gate -> z, then the 1x1s);
 * d_z is a caller-provided (B,L,D) workspace.  Shapes: R % 256 == 0, S % 256 == 0, D % 128 == 0 (wn_tc_supported).
 * Weights are packed K-major and pre-split: d_wa [2][2D][k*R] (rows in 256-wide tiles: 128 filter channels then the
 * same 128 gate channels; column j*R+r = tap j of input channel r; [0]=hi, [1]=lo), d_ba [2D] in the same row order,
 * d_wb [2][R+S][D] (residual rows then skip rows), d_bb [R+S]. */
int wn_tc_supported(int R, int D, int S, int k);
int wn_tc_pack_block_weights(const float* d_wf, const float* d_wg, const float* d_bf, const float* d_bg,
                             const float* d_wr, const float* d_ws, const float* d_br, const float* d_bs,
                             int R, int D, int S, int k, float* d_wa, float* d_ba, float* d_wb, float* d_bb, void* stream);
typedef struct wn_tc_block_args {
    const float* d_h_in; float* d_h_out; float* d_skip; float* d_z;
    const float* d_wa; const float* d_ba; const float* d_wb; const float* d_bb;
    int B, L, R, D, S, k, dilation;
    int in_start, out_start, skip_start, skip_init;
    float* d_fg_save;
    int fast_tf32;      /* precision mode.  0 = 3xTF32 (hi/lo tf32 split of both operands, ~6e-7 on the logits after 50
                         * layers).  2 = bf16 pairs (hi/lo bf16 split, the same three products at twice the MMA rate,
                         * ~3e-6 on the logits after 50 layers; d_wa / d_wb must then be the arrays made by
                         * wn_tc_convert_weights_bf16).  1 = single TF32 pass: ~1e-3 on the logits -- OUTSIDE the 1e-4
                         * parity bar; opt-in, reported separately by bench.py */
} wn_tc_block_args;
int wn_tc_block_fwd(const wn_tc_block_args* a, void* stream);
/* Re-split a packed fp32 pair array [2][n_per_half] (hi | lo, as written by wn_tc_pack_block_weights /
 * wn_tc_pack_block_bwd_weights) into bf16 pairs [2][n_per_half] for precision mode 2: x = hi + lo exactly,
 * out_hi = bf16(x), out_lo = bf16(x - out_hi).  d_out: 4 * n_per_half bytes. */
int wn_tc_convert_weights_bf16(const float* d_pairs, void* d_out, long long n_per_half, void* stream);
/* Debug aid (WN_TC_TRACE=1): per-stage clock64 stamps of CTA 0 of the most recent tensor-core launch, 8 per stage:
 * producer before/after the empty wait, MMA warp before/after the operand wait and after issue, splitter start/end
 * (first splitter warp), end of the last splitter warp. */
int wn_tc_read_trace(long long* host_out, int n);

/* ---------------------------------------------------------------- (T) the block as ONE tensor-core launch (round 2 default)
 * Same mathematics as wn_block_fwd (reference wavenet_model.py:142-165) for R = D = S = 256, k = 2 (wn_tb_supported), with
 * fp32-class accuracy from bf16 (hi, lo) operand pairs on tcgen05.mma cta_group::2 (fp32 accumulation in tensor memory); the
 * gated activation z never leaves the SM.  Activations use the CHUNKED PAIR LAYOUT: a (B, L, C) activation is stored as
 *     bf16 [b][plane: 0 = hi, 1 = lo][c / 8][t][c % 8]         x = hi + lo, hi = bf16(x), lo = bf16(x - hi)
 * (the same number of bytes as fp32 frames), and skip as fp32 [b][c / 4][t - skip_start][c % 4].  wn_pair_from_frames /
 * wn_frames_from_pair / wn_frames_from_chunks4 convert to and from the frames layout of the other entry points.
 * Weights: wn_tb_pack_all_weights writes every layer's pre-split, pre-tiled image (wn_tb_weight_bytes_per_layer(channels,
 * precision) bytes each) into ONE array d_w_all [n_layers][bytes_per_layer] and the bias vectors [bf | bg | br | bs].
 * wn_tb_start_index_* is start_conv on class indices (wavenet_model.py:65-68,127) writing that layout; *d_err (optional) is
 * set to 1 when an index is outside [0, classes) -- the reference's one-hot scatter would raise there. */
int    wn_tb_supported(int R, int D, int S, int k);                  /* R = D = S in {256, 512}, k = 2                        */
int    wn_tb_precision_supported(int channels, int precision);       /* pairs: 256 channels; single-pass bf16: 256 or 512   */
size_t wn_tb_weight_bytes_per_layer(int channels, int precision);
/* all layers in one launch: d_ptrs is a DEVICE table [n_layers][8] of {wf, wg, bf, bg, wr, ws, br, bs} (biases may be 0);
 * d_bias_all receives [n_layers][4 * channels] = [bf | bg | br | bs] */
int    wn_tb_pack_all_weights(const float* const* d_ptrs, int n_layers, int channels, int precision, void* d_w_all,
                              float* d_bias_all, void* stream);
int    wn_tb_start_index_u8(const uint8_t* d_idx, const float* d_w_t, const float* d_b_p, void* d_h_pair,
                            int B, int classes, int L, int R, int* d_err, void* stream);
int    wn_tb_start_index_i64(const int64_t* d_idx, const float* d_w_t, const float* d_b_p, void* d_h_pair,
                             int B, int classes, int L, int R, int* d_err, void* stream);
int    wn_pair_from_frames(const float* d_frames, void* d_pair, int B, int L, int C, int t_begin, void* stream);
int    wn_frames_from_pair(const void* d_pair, float* d_frames, int B, int L, int C, int t_begin, void* stream);
/* frames [t_first, t_first + n) of a chunked fp32 tensor (B, C/4, T, 4) -> (B, n, C) */
int    wn_frames_from_chunks4(const float* d_chunked, float* d_frames, int B, int T, int C, int t_first, int n, void* stream);
typedef struct wn_tb_block_args {
    const void* d_h_in; void* d_h_out;     /* chunked pairs (B, 2, channels/8, L, 8) bf16                           */
    float* d_skip;                         /* chunked (B, channels/4, L - skip_start, 4) fp32                       */
    const void* d_w_all; const float* d_bias4;   /* all layers' packed weights; THIS layer's biases [4 * channels] */
    int layer, n_layers;
    int channels, precision;               /* 256 / 512; WN_PREC_*                                                  */
    int B, L, dilation;
    int in_start, out_start, skip_start, skip_init;
    float* d_fg_save;                      /* optional chunked (B, 2*channels/4, L, 4) fp32: tanh | sigmoid outputs (for the backward) */
} wn_tb_block_args;
int    wn_tb_block_fwd(const wn_tb_block_args* a, void* stream);

/* ALL residual blocks of a forward in ONE persistent launch: the (layer, 256-frame item) list is dealt round-robin to the CTA
 * pairs, an item waits for the previous layer's items that wrote the frames it reads (device-side flags), so there is no launch
 * gap and no idle tail between layers.  h_ptrs: HOST array [n_layers + 1] of device pair tensors, layer i reads h_ptrs[i] and
 * writes h_ptrs[i+1]; buffers may repeat with period >= 3 (never h_ptrs[i+1] == h_ptrs[i] or h_ptrs[i-1]).  d_bias_all:
 * [n_layers][4*channels]; d_fg_all: optional (n_layers, B, 2*channels/4, L, 4); d_desc: n_layers * wn_tb_stack_desc_bytes()
 * bytes of 128-byte aligned device scratch; d_flags: (wn_tb_stack_items(...) + n_layers) uint32 of device scratch.
 * dilations / in_start / out_start: HOST arrays [n_layers]. */
size_t    wn_tb_stack_desc_bytes(void);
long long wn_tb_stack_items(int n_layers, int B, int L, const int* out_start);
typedef struct wn_tb_stack_args {
    const void* const* h_ptrs;
    float* d_skip; const void* d_w_all; const float* d_bias_all; float* d_fg_all;
    void* d_desc; unsigned* d_flags;
    int n_layers, channels, precision, B, L, skip_start;
    const int* dilations; const int* in_start; const int* out_start;
} wn_tb_stack_args;
int       wn_tb_stack_fwd(const wn_tb_stack_args* a, void* stream);

/* Backward of the same block on the same layout (tcgen05 cta_group::2, bf16 pairs), replacing autograd's backward through
 * wavenet_model.py:142-165.  Frame-range arguments are those of wn_block_bwd_args.  Buffers: d_dh_out (B,2,32,L,8) pair or
 * NULL (last layer), d_dskip (B,2,32,L-ds_start,8) pair on its own frame axis, d_fg the forward's d_fg_save, outputs d_dfg
 * (B,2,64,L,8) pair [dF chunks 0..31 | dG chunks 32..63], d_z (B,2,32,L,8) pair (recomputed tanh*sigmoid), d_dh_in pair.
 * d_wb_all: [n_layers][wn_tb_bwd_weight_bytes_per_layer(channels, precision)] images written by wn_tb_pack_all_bwd_weights. */
size_t wn_tb_bwd_weight_bytes_per_layer(int channels, int precision);
int    wn_tb_pack_all_bwd_weights(const float* const* d_ptrs, int n_layers, int channels, int precision, void* d_wb_all,
                                  void* stream);
typedef struct wn_tb_bwd_args {
    const void* d_dh_out; const void* d_dskip; const float* d_fg;
    void* d_dfg; void* d_z; void* d_dh_in;
    const void* d_wb_all;
    int layer, n_layers;
    int channels, precision;
    int B, L, dilation;
    int in_start, out_start;
    int gs_out, ds_start, gz, gs_in;
} wn_tb_bwd_args;
int    wn_tb_block_bwd_data(const wn_tb_bwd_args* a, void* stream);
/* All weight gradients of one block in one launch (+ a deterministic reduction): the contraction over frames reads the
 * chunked tiles as MN-major tcgen05 operands.  Outputs are the parameter-shaped tensors: d_gws (S,D,1), d_gwr (R,D,1),
 * d_gwf / d_gwg (D,R,2).  id_start: first frame where dh_out flows straight into dh_in (= max(out_start, gs_out)).
 * d_work: wn_tb_wgrad_workspace_bytes() bytes. */
size_t wn_tb_wgrad_workspace_bytes(void);
typedef struct wn_tb_wgrad_args {
    const void* d_dskip; const void* d_dh_out; const void* d_dfg; const void* d_z; const void* d_h_in;
    float* d_gws; float* d_gwr; float* d_gwf; float* d_gwg; float* d_work;
    int channels, precision;
    int B, L, dilation;
    int in_start, ds_start, id_start, gz;
} wn_tb_wgrad_args;
int    wn_tb_wgrad(const wn_tb_wgrad_args* a, void* stream);

/* ---------------------------------------------------------------- (T) head
 * replaces relu -> end_conv_1 -> relu -> end_conv_2 (wavenet_model.py:167-169) and forward()'s
 * slice/transpose/view (:191-196): logits (B*out_len, classes) for the LAST out_len frames only.
 * d_skip is (B, L-skip_start, S); requires out_len <= L-skip_start (the reference raises on view otherwise). */
typedef struct wn_head_args {
    const float* d_skip; float* d_logits;
    const float* d_w1_t; const float* d_b1; const float* d_w2_t; const float* d_b2;
    int B, L, S, E, classes, skip_start, out_len;
    int mode;
} wn_head_args;
int wn_head_fwd(const wn_head_args* a, void* stream);

/* ---------------------------------------------------------------- (T) backward, data gradients
 * replace autograd's backward through the layer loop (the reference calls loss.backward(), wavenet_training.py:71).
 * Gradient buffers use the frames layout; each carries a first valid frame, left of which it is structurally zero
 * and is neither read nor written:
 *   gs_out   first frame where d_dh_out may be non-zero (L if the block output is unused, as for the last layer)
 *   ds_start first frame of d_dskip, which is (B, L-ds_start, S)       (= L - output_length)
 *   gz       first frame for which dz / d_dfg / d_z are produced       (>= out_start)
 *   gs_in    first frame for which d_dh_in is produced                 (>= in_start)
 * d_wrs_rows: [(R+S)][wn_n2p(D)] rows of residual_conv.weight then skip_conv.weight (zero padded columns);
 * d_wfg_bwd : [k*2D][wn_n2p(R)], row j*2D+n holds [filter;gate].weight[n, :, j].
 * Weight gradients are plain GEMMs over the produced buffers (dWr = dh_out^T z, dWs = dskip^T z,
 * dW{f,g}[:,:,j] = d{F,G}^T h_in(t-(k-1-j)d), biases = column sums) and are left to the caller. */
typedef struct wn_block_bwd_args {
    const float* d_dh_out; const float* d_dskip; const float* d_fg;
    float* d_dfg; float* d_z; float* d_dh_in;
    const float* d_wrs_rows; const float* d_wfg_bwd;
    int B, L, R, D, S, k, dilation;
    int in_start, out_start;
    int gs_out, ds_start, gz, gs_in;
} wn_block_bwd_args;
int wn_block_bwd_data(const wn_block_bwd_args* a, void* stream);

/* The same two data-gradient GEMMs on the tensor cores (tcgen05, 3xTF32), for R % 256 == 0, S % 256 == 0, D % 256 == 0:
 * weights packed K-major and pre-split by wn_tc_pack_block_bwd_weights into d_wdz [2][D][R+S] (row c: residual_conv
 * column c then skip_conv column c) and d_wdh [2][R][k*2D] (row r, column j*2D+n: [filter;gate].weight[n][r][j]).
 * d_wrs_rows / d_wfg_bwd of the args are ignored. */
int wn_tc_bwd_supported(int R, int D, int S, int k);
int wn_tc_pack_block_bwd_weights(const float* d_wf, const float* d_wg, const float* d_wr, const float* d_ws,
                                 int R, int D, int S, int k, float* d_wdz, float* d_wdh, void* stream);
int wn_tc_block_bwd_data(const wn_block_bwd_args* a, const float* d_wdz, const float* d_wdh, void* stream);
/* the same with a precision mode: 0 = 3xTF32 (fp32 pair arrays), 2 = bf16 pairs (arrays from wn_tc_convert_weights_bf16) */
int wn_tc_block_bwd_data_prec(const wn_block_bwd_args* a, const void* d_wdz, const void* d_wdh, int precision, void* stream);

/* head: given d_dlogits (B*out_len, classes) and the saved skip sum (B, L-skip_start, S) produce
 * d_y1 (B*out_len, E) = relu(W1 relu(skip)+b1) (recomputed), d_dy1 (B*out_len, E) and d_dskip (B, out_len, S).
 * d_w1_t/d_b1: end_conv_1 packed by wn_pack_1x1_weights; d_w2_rows [classes][wn_n2p(E)] = end_conv_2.weight rows;
 * d_w1_rows [E][wn_n2p(S)] = end_conv_1.weight rows. */
typedef struct wn_head_bwd_args {
    const float* d_dlogits; const float* d_skip;
    float* d_y1; float* d_dy1; float* d_dskip;
    const float* d_w1_t; const float* d_b1; const float* d_w2_rows; const float* d_w1_rows;
    int B, L, S, E, classes, skip_start, out_len;
} wn_head_bwd_args;
int wn_head_bwd_data(const wn_head_bwd_args* a, void* stream);

/* weight gradient of one convolution tap -- what autograd's conv1d backward-weight computes for filter/gate
 * (wavenet_model.py:145-151), residual/skip (:156,:164) and the head convolutions (:167-169):
 *     dw[n*dw_n_stride + c*dw_c_stride] = sum_{b<B} sum_{t<rows} g[b*g_seq_stride + t*ldg + n] * x[b*x_seq_stride + t*ldx + c]
 * for n < N, c < C (overwrites, exact fp32, deterministic).  d_g / d_x point at the first paired frame of sequence 0
 * (a tap shift is a pointer offset); strides are in floats, so the result can be written straight into column j of
 * an (out, in, k) weight-gradient tensor (dw_n_stride = in*k, dw_c_stride = k).  d_work: wn_wgrad_workspace_bytes(N, C)
 * bytes of scratch for the split-frames partial sums.  rows == 0 writes zeros. */
typedef struct wn_wgrad_args {
    const float* d_g; const float* d_x; float* d_dw; float* d_work;
    long long g_seq_stride, x_seq_stride, dw_n_stride, dw_c_stride;
    int ldg, ldx, B, rows, N, C;
} wn_wgrad_args;
size_t wn_wgrad_workspace_bytes(int N, int C);
int wn_wgrad(const wn_wgrad_args* a, void* stream);
/* The same contraction on the tensor cores (tcgen05 kind::f16 on bf16 hi/lo pairs, fp32 accumulation; the operand
 * tiles are transposed to K-major while they are split): C == 256, N % 128 == 0, rows >= 1, ldg / ldx / sequence
 * strides multiples of 4 floats, d_g / d_x 16-byte aligned.  Same argument block and workspace as wn_wgrad. */
int wn_tc_wgrad_supported(int N, int C);
int wn_tc_wgrad(const wn_wgrad_args* a, void* stream);

/* ---------------------------------------------------------------- (T) the rest of a training step
 * What WavenetTrainer.train does around model(x) (reference wavenet_training.py:64-76), as native launches:
 * wn_ce_fwd_bwd: F.cross_entropy(output, target) (:69) AND its gradient in one pass: *d_loss = mean_i(logsumexp(x_i) -
 *   x_i[target_i]), d_dlogits = (softmax(x_i) - onehot(target_i)) / N (may not alias d_logits); deterministic; *d_err
 *   (optional) is set when a target is outside [0, C).  d_work: wn_ce_workspace_bytes().  C <= 1024.
 * wn_adam_step: torch.optim.Adam's update (the reference's default optimizer, :24) of every tensor in one launch.  d_segs
 *   is a DEVICE array of segments, d_chunks a DEVICE array of (segment, chunk-of-4096) int pairs covering them.
 * wn_scatter_rows: start_conv gradient for index input: table (classes, R) = sum over frames t >= t_begin of dh[b][t][:]
 *   into row idx[b][t] (idx uint8 or int64, (B, L)), optionally also transposed into d_out_t (R, classes) = the layout of
 *   start_conv.weight; wn_colsum: out[c] = sum_r x[r][c] (bias gradients), d_work
 *   wn_colsum_workspace_bytes(rows, C); wn_relu_copy: y = max(x, 0) (the head's relu(skip) operand of a weight gradient). */
size_t wn_ce_workspace_bytes(void);
int    wn_ce_fwd_bwd(const float* d_logits, const int64_t* d_target, float* d_dlogits, float* d_loss, float* d_work, int* d_err,
                     int N, int C, void* stream);
typedef struct wn_adam_seg { float* p; const float* g; float* m; float* v; long long n; } wn_adam_seg;
int    wn_adam_step(const wn_adam_seg* d_segs, const int* d_chunks, int n_chunks, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, void* stream);
int    wn_scatter_rows(const void* d_idx, int idx_is_u8, const float* d_dh, float* d_table, float* d_out_t, int B, int L, int R,
                       int classes, int t_begin, void* stream);
size_t wn_colsum_workspace_bytes(long long rows, int C);
int    wn_colsum(const float* d_x, float* d_out, float* d_work, long long rows, int C, int ld, void* stream);
int    wn_relu_copy(const float* d_x, float* d_y, long long n, void* stream);
/* x *= *d_scale in place (d_scale: one float on the device -- the gradient autograd passes into the loss node,
 * wavenet_training.py:71 `loss.backward()`); no pass over x when the scalar is 1.  n % 4 == 0. */
int    wn_scale_by(float* d_x, long long n, const float* d_scale, void* stream);

/* ---------------------------------------------------------------- (G) Fast-WaveNet sampler
 * replaces WaveNetModel.generate_fast's warm-up and sampling loops (wavenet_model.py:250-302) together with
 * DilatedQueue.enqueue/dequeue/reset (wavenet_modules.py:55-77): ONE persistent cooperative kernel runs
 * `n_evals` network evaluations without returning to the host; the per-layer ring buffers live in d_rings.
 *
 * Weights are the reference's own parameter tensors, UNPACKED (state_dict layout); bias pointers may be NULL.
 * n_streams independent streams share the weights (the reference has a single stream, wavenet_model.py:179).
 * Run through the SAME kernel (wn_gen_set_mode), stream s of a multi-stream run equals a single-stream run with the
 * same inputs bit for bit; different kernels split the dot products differently (rounding-level differences). */
typedef struct wn_gen_weights {
    const float* d_start_w; const float* d_start_b;            /* (R,classes,1), (R) */
    const float* const* d_wf; const float* const* d_bf;        /* HOST arrays [n_layers] of device pointers */
    const float* const* d_wg; const float* const* d_bg;        /* (D,R,k), (D)   */
    const float* const* d_wr; const float* const* d_br;        /* (R,D,1), (R)   */
    const float* const* d_ws; const float* const* d_bs;        /* (S,D,1), (S)   */
    const float* d_end1_w; const float* d_end1_b;              /* (E,S,1), (E)   */
    const float* d_end2_w; const float* d_end2_b;              /* (classes,E,1)  */
} wn_gen_weights;

typedef struct wn_gen_shape {
    int n_layers, k, R, D, S, E, classes, n_streams;
    const int* dilations;                                      /* HOST array [n_layers] */
} wn_gen_shape;

/* bytes the caller must provide: rings (all layers, all streams) and scratch (exchange vectors, barrier) */
int wn_gen_workspace_bytes(const wn_gen_shape* s, size_t* ring_bytes, size_t* scratch_bytes);

typedef struct wn_gen_handle wn_gen_handle;
int wn_gen_create(const wn_gen_shape* s, const wn_gen_weights* w, float* d_rings, void* d_scratch,
                  wn_gen_handle** out);
/* zero the rings and the time counter (DilatedQueue.reset, wavenet_modules.py:74-77) */
int wn_gen_reset(wn_gen_handle* h, void* stream);
/* Run evaluations [t0, t0+n_evals) of a schedule with n_given given samples per stream:
 *   evaluation e takes as input  d_first[s*n_given + e]            if e <  n_given
 *                                d_forced[s*n_samples + e-n_given] if d_forced != NULL   (teacher forcing)
 *                                the index chosen at evaluation e-1 otherwise,
 *   and, when e >= n_given-1, chooses sample i = e-(n_given-1): argmax of (logits - regularize*(c-classes/2)^2)
 *   if temperature <= 0 (wavenet_model.py:290-294), else inverse-CDF sampling of softmax(./temperature) with the
 *   float64 uniform d_uniforms[s*n_samples + i] exactly as numpy.random.choice does (wavenet_model.py:282-289).
 *   d_out_idx (n_streams, n_samples) int32; d_out_logits optional (n_streams, n_samples, classes) fp32 holding
 *   logits - regularizer.  t0 must continue where the previous call stopped (0 after wn_gen_reset).          */
typedef struct wn_gen_run_args {
    const int32_t* d_first; int n_given;
    const int32_t* d_forced; const double* d_uniforms;
    int32_t* d_out_idx; float* d_out_logits;
    int n_samples;            /* row pitch of forced / uniforms / out_idx                  */
    int t0, n_evals;
    float temperature, regularize;
} wn_gen_run_args;
int wn_gen_run(wn_gen_handle* h, const wn_gen_run_args* a, void* stream);
int wn_gen_destroy(wn_gen_handle* h);
/* Which sampler kernel runs (all implement the same schedule; call right after wn_gen_reset):
 *   0  auto: 256-wide k = 2 nets -> the tensor-core cluster kernel 6, for any number of streams (measured: one stream
 *            107.7 us/sample against 150.7 for kernel 3; 64 streams 382 k samples/s against 23 k for kernel 4);
 *            other nets: one stream -> the single-stream L2 kernel 3, several streams -> one cluster per stream
 *            (kernel 4) where it applies, otherwise the generic kernel
 *   1  atomic grid barrier between stages (the simple reference kernel)
 *   2  generic flag-in-data exchange through L2 (any shape, any number of streams)
 *   3  single-stream L2 kernel with register-free cooperative polling (k = 2, power-of-two row split)
 *   4  cluster kernel: one 16-CTA thread-block cluster per stream, exchange through distributed shared memory
 *   5  single-stream two-level exchange: kernel 3's grid (64 CTAs x 4 rows) as 4 clusters of 16; values go to the 16 CTAs
 *      of the producer's cluster through distributed shared memory and reach the other clusters through ONE L2 poller per
 *      (cluster, producer) that forwards them by DSMEM (256-wide nets: R = D = S = E = classes = 256).  Measured 2x slower
 *      than kernel 3 (a 16-CTA DSMEM all-to-all costs as much as the L2 one it replaces): selectable, never the default
 *   6  tensor-core cluster kernel: up to 8 streams per thread-block cluster (256-wide nets: R = D = S = E = classes = 256,
 *      k = 2).  The weights of a stage enter shared memory once per 8 streams, as bf16 hi/lo pairs pre-split into MMA
 *      fragment order at wn_gen_reset (wn_gen_workspace_bytes includes the images; wn_gen_weights_changed after in-place
 *      weight updates); dot products are mma.sync m16n8k16 with three MMAs per product (fp32-class: ~1e-6 on the
 *      logits); the exchange is one 512-byte st.async.v4 block per destination CTA, credited to an mbarrier there.
 *      Clusters of 16 CTAs while all of them are co-resident (<= 7 clusters = 56 streams on a B200), else clusters of 8
 *      CTAs that own two 16-channel slices each (<= 15 clusters = 120 streams per wave)
 * Kernels 2, 3 and 5 sum in the same order (bit-identical results); kernel 4 splits rows differently (rounding-level
 * differences). */
int wn_gen_set_mode(wn_gen_handle* h, int mode);
/* The parameter tensors given to wn_gen_create were written in place (an optimizer step): kernels 1-5 read them on every
 * launch and need nothing; kernel 6 keeps pre-split copies, which the next wn_gen_reset rebuilds after this call. */
int wn_gen_weights_changed(wn_gen_handle* h);
/* Synchronise the stream and report whether a launch aborted (a CTA waited > ~3 s for a tag): 0 = fine. */
int wn_gen_check(wn_gen_handle* h, void* stream);
/* Debug aid: with WN_GEN_TRACE=1 in the environment at wn_gen_create, CTA 0 stamps clock64() at 8 points of every layer
 * of the LAST evaluation of a launch (single-stream kernel only); this copies the first n stamps to the host. */
int wn_gen_read_trace(wn_gen_handle* h, long long* host_out, int n, void* stream);
/* Which kernel wn_gen_run launches in the handle's current mode: the number (1-6) documented at wn_gen_set_mode. */
int wn_gen_kernel_id(const wn_gen_handle* h);
/* how wn_gen_run launches: grid size, block size, dependent exchange stages per evaluation */
int wn_gen_launch_info(const wn_gen_handle* h, int* grid, int* block, int* barriers_per_eval);

#ifdef __cplusplus
}
#endif
#endif /* WAVENET_B200_H */
