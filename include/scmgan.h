/*
 * scmgan.h - C ABI of libscmgan.so: hand-written sm_100a kernels behind the scm-gan world-model training step.
 *
 * The reference (LilJing/scm-gan) is pure PyTorch and has no FFI of its own; each entry point below names the
 * reference call site whose library kernels (cuDNN / cuBLAS / ATen) it replaces.  The host side that binds
 * these (scm_gan_b200/_lib.py, ctypes) mirrors reference models.py / spectral_normalization.py one level up.
 *
 * Conventions
 *  - every pointer is a CUDA device pointer unless the name ends in _host; buffers are owned by the caller
 *    (torch's caching allocator); the library allocates nothing persistent and never synchronises;
 *  - every call enqueues work on `stream` and is CUDA-graph-capture safe;
 *  - return value: 0 = ok, SCMGAN_EINVAL (-1) bad argument, SCMGAN_EUNSUPPORTED (-2) shape not supported,
 *    SCMGAN_ECUDA (-3) CUDA runtime/driver error; scmgan_last_error() returns a thread-local message;
 *  - "plane": bf16 activation tensor [B][H+2][W+2][Cs] (NHWC + 1-pixel halo, Cs multiple of 8).  The halo
 *    holds zeros (zero padding: Encoder/Decoder/RewardPredictor) or the wrapped border (the legacy circular
 *    pad-1 of Transition, reference models.py:51-56 under torch<=1.4 semantics).
 */
#ifndef SCMGAN_H_
#define SCMGAN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCMGAN_OK 0
#define SCMGAN_EINVAL (-1)
#define SCMGAN_EUNSUPPORTED (-2)
#define SCMGAN_ECUDA (-3)

#define SCMGAN_ACT_NONE 0
#define SCMGAN_ACT_LRELU 1
#define SCMGAN_ACT_SIGMOID 2

/* 16-bit storage formats of planes and packed weights.  Forward activations and weights are fp16 by default (11-bit
 * significand: parameter gradients then stay within 1e-2 of the fp32 reference, profiles/r02_grad_precision_*.json),
 * gradient planes are bf16 (fp32 exponent range).  One tcgen05.mma (kind::f16) needs both operands in the SAME format
 * (mixing faults with an illegal instruction on B200), so: forward = fp16 x fp16; dgrad = bf16 gradient x bf16-packed
 * weights; wgrad = bf16 gradient x fp16 activation, the activation tile being rewritten to bf16 in shared memory by
 * the kernel (scm_gan_b200/csrc/conv_wgrad.cuh). */
#define SCMGAN_FMT_BF16 0
#define SCMGAN_FMT_F16 1

typedef void* scmgan_stream_t; /* cudaStream_t */

int scmgan_version(void);
const char* scmgan_last_error(void);
int scmgan_num_sms(void);
/* number of kernels this library has launched so far in this process (diagnostics: bench.py "gpu_launches") */
long long scmgan_launch_count(void);

/* fp32 NCHW tensor (batch stride src_bstride elements, channel stride H*W) -> plane channels
 * [c_off, c_off+c_pad), zero-filling channels >= C; halo = wrap (1) or zeros (0).
 * sig (optional, dense fp32 [B][C][H][W]): multiply by sig*(1-sig) - the backward of the output sigmoid
 * (reference models.py:103,154) fused into the packing of the incoming gradient.
 * Replaces x.view / torch.cat / F.pad(mode='circular') copies: reference models.py:69-73, 143, 76-103. */
int scmgan_pack_nchw(const float* src, long long src_bstride, int C, int B, int H, int W, void* dst_plane, int Cs,
                     int c_off, int c_pad, int wrap, const float* sig, int fmt, scmgan_stream_t stream);

/* CoordConv coordinate channels (reference coordconv.py:10-14; interface-only layer): plane channel c_off of interior
 * pixel (h, w) <- -1 + 2w/W and channel c_off+1 <- -1 + 2h/H, for every sample.  Run after scmgan_pack_nchw has
 * filled (zeroed) the channel window; replaces the arange/repeat/cat of the reference. */
int scmgan_pack_coords(void* dst_plane, int Cs, int c_off, int B, int H, int W, int fmt, scmgan_stream_t stream);

/* Weight packing fp32 parameter -> bf16 [9][n_pad][k_pad] GEMM operand, optionally divided by *sigma
 * (the `w / sigma` of reference spectral_normalization.py:35).
 *   out[tap][n][k] = w[n*s_n + (k+k_src_off)*s_k + (flip ? 8-tap : tap)] / sigma
 * out_ld (elements, 0 = k_pad) is the row pitch of `out`: with out_ld > k_pad the job fills a K window of a wider
 * operand [9][n_pad][out_ld], so that the gradients of two convolutions reading the same activation (the skip
 * connections of reference models.py:95,101) become ONE dgrad GEMM over concatenated K. */
typedef struct {
    const float* w;
    void* out;
    const float* sigma; /* device scalar or NULL */
    int n_pad, k_pad, n_valid, k_valid;
    long long s_n, s_k;
    int k_src_off;
    int flip;
    int out_ld;
    int fmt; /* SCMGAN_FMT_* of `out` */
} scmgan_pack_job;
int scmgan_pack_weights(int count, const scmgan_pack_job* jobs_host, scmgan_stream_t stream);

/* 3x3 stride-1 same-size convolution over a plane as a tcgen05 implicit GEMM, with fused epilogue.
 * Forward of nn.Conv2d / nn.ConvTranspose2d(stride 1, pad 1) (reference models.py:76-103, 145-154, 274-279)
 * and, with flipped/transposed packed weights, their data gradient (cuDNN dgrad under main.py:285).
 * Epilogue: y = acc*scale + bias[n] + sample_bias[b][n]; y += add; y = act(y); y *= lrelu'(gate);
 *   -> bf16 plane `out` (interior + wrapped halo, or zero halo) and/or fp32 NCHW `out_f32`
 *   -> optional Bernoulli head: sample_out = (uniforms < y), or (philox(rng_state) < y), or (y > 0.5)
 *      (reference models.py:30-40, 107-112). */
typedef struct {
    int B, H, W;
    const void* x; /* input plane */
    int x_cs, x_c_off, cin; /* channel stride, first channel, channel count (multiple of 16) */
    const void* w;          /* packed weights [9][n][cin] */
    int n;                  /* GEMM N: multiple of 16, <= 256 */
    float scale;
    const float* bias;        /* [n] or NULL */
    const float* sample_bias; /* [B][n] or NULL */
    int act;
    float slope;
    void* out; /* bf16 plane or NULL */
    int out_cs, out_c_off, wrap;
    const void* add; /* bf16 plane or NULL */
    int add_cs, add_c_off;
    const void* gate; /* bf16 plane or NULL */
    int gate_cs, gate_c_off;
    float* out_f32; /* [B][n_valid][H][W] or NULL */
    int n_valid;
    float* sample_out;     /* [B][n_valid][H][W] or NULL */
    const float* uniforms; /* [B][n_valid][H][W] or NULL */
    unsigned long long* rng_state; /* device {seed, offset}: with sample_out and uniforms == NULL the Bernoulli draw uses an
                                      in-kernel Philox4x32-10 stream and the offset is advanced; NULL => threshold 0.5 */
    int bias_n; /* number of valid entries of `bias` (channels beyond it get 0); 0 = n.  Lets a layer with 3 or 12
                   real output channels pass its own bias vector although n is padded to 16. */
    int x_fmt, w_fmt, out_fmt; /* SCMGAN_FMT_* of the input plane, the packed weights and the output plane */
    const float* sample_scale; /* [B] or NULL: y = acc*scale*sample_scale[b] + ...  Samples of several spectral-norm calls
                                  (different sigma: the main rollout step and the counterfactual rollouts of reference
                                  main.py:242-283) then share one GEMM with weights packed for the first call's sigma. */
    int coord_c1;              /* CoordConv (reference coordconv.py:5-15): 0 = off; otherwise 1 + the index (even, inside the
                                  input window) of the x-coordinate channel, the y coordinate follows it.  The two channels
                                  are GENERATED while the producer stages the im2col tile (x = -1 + 2w/W, y = -1 + 2h/H,
                                  zero in the padding halo); the plane holds zeros there.  Needs cin == 16, W <= 69 and a
                                  zero-padded plane; wider layers materialise the channels with scmgan_pack_coords. */
    int weights_stable;        /* 1: the caller guarantees that `w` was last written at least two kernel launches before this
                                  call in stream order (e.g. data-gradient operands packed during the forward pass): the
                                  CTA-pair kernel may then load its resident weights before griddepcontrol.wait, i.e. while
                                  the preceding kernel drains.  0: the library decides from its own launch bookkeeping. */
} scmgan_conv_desc;
int scmgan_conv3x3_fwd(const scmgan_conv_desc* desc_host, scmgan_stream_t stream);
int scmgan_conv3x3_dgrad(const scmgan_conv_desc* desc_host, scmgan_stream_t stream);

/* One directional sweep of CSRN (reference spatial_recurrent.py:61-114; interface-only layer: imported by models.py:14,
 * never instantiated by main.py).  Lines are visited in order (or reversed); per line: single-step bias-free GRU on the
 * line's n pixels with the hidden state handed over from the previous line, context = GRU output, next hidden state =
 * tanh(Conv1d_k3,p1(output) + b).  x, ctx, dx and dctx share the strides xs_* (element units) over [B][C][line][pixel]:
 * a row sweep of a dense [B,C,H,W] tensor uses xs_line = W, xs_pix = 1, L = H, n = W; a column sweep xs_line = 1,
 * xs_pix = W, L = W, n = H.  `states` [B][L][n][C] receives the hidden state entering every line (needed by bwd).
 * bwd: dctx and ctx (the forward output) in, dx out, and per-sample parameter-gradient partials
 * dparams [B][3C*C (dW_ih) + 3C*C (dW_hh) + 3C*C (dconv_w, [C][C][3]) + C (dconv_b)] which the caller zeroes
 * beforehand and sums over B afterwards (deterministic).  Limits: 12*n*C*4 bytes of shared memory (<= 227 KB). */
typedef struct {
    const float* x;
    long long xs_b, xs_c, xs_line, xs_pix;
    int B, C, L, n;
    int reverse;
    const float* w_ih;   /* [3C][C] GRU weight_ih_l0 (gates r, z, n) */
    const float* w_hh;   /* [3C][C] GRU weight_hh_l0 */
    const float* conv_w; /* [C][C][3] */
    const float* conv_b; /* [C] */
    float* ctx;          /* fwd: output.  bwd: forward output (read) */
    float* states;       /* [B][L][n][C] */
    const float* dctx;   /* bwd */
    float* dx;           /* bwd */
    float* dparams;      /* bwd */
} scmgan_csrn_sweep_desc;
int scmgan_gru_conv_sweep_fwd(const scmgan_csrn_sweep_desc* desc_host, scmgan_stream_t stream);
int scmgan_gru_conv_sweep_bwd(const scmgan_csrn_sweep_desc* desc_host, scmgan_stream_t stream);

/* out[i] = uniform in (0,1) number i of the Philox4x32-10 stream {seed, offset} held in rng_state (element i uses
 * counter offset + i/4, lane i%4: the very numbers the in-kernel Bernoulli head of scmgan_conv3x3_fwd draws for
 * element i), then offset += ceil(n/4).  Feeds `uniforms` of the Transition's last conv: generating the stream in
 * its own pass (all four outputs of every Philox block used, full-chip parallelism) is ~10x cheaper than inside the
 * conv epilogue.  Replaces torch.rand_like of DifferentiableBernoulliSampler, reference models.py:27-31. */
int scmgan_philox_uniform(float* out, long long n, unsigned long long* rng_state, scmgan_stream_t stream);

/* Weight gradient of the two coordinate channels of a CoordConv whose coordinates were generated in the tile
 * (scmgan_conv_desc::coord_c1): g[co*g_s_co + (coord_c + c)*g_s_ci + tap] = sum_{b,h,w} dy[b][co][h][w] * coord_c(h+ky-1,
 * w+kx-1), zero outside the image.  dy is the fp32 NCHW incoming gradient.  The tensor-core weight-gradient kernels read
 * the input plane, which holds zeros in those two channels. */
int scmgan_coord_wgrad(const float* dy, int B, int Co, int H, int W, float* g, long long g_s_co, long long g_s_ci,
                       int coord_c, scmgan_stream_t stream);

/* Device-side replay buffer sampler: the reference's get_trajectories (envs/minipacman.py:122-164; same contract in
 * envs/betterpong.py, gridworld.py, ...) without leaving the GPU.  Episodes occupy fixed slots in HBM; the first
 * *n_filled slots are valid, ep_len[slot] (>= 4) is the episode length.  Each batch row concatenates clips
 * (random episode, random start in [0, len-4] or 0, up to len-1) until Hn timesteps are filled; the last step of every
 * clip - also of the final, truncated one - is flagged done, as in the reference.  Randomness: Philox4x32-10 stream
 * {seed, offset} in rng_state; clip k of row b uses counter offset + b*Hn + k (lanes 0, 1); offset += B*Hn afterwards
 * (second launch), so the call is CUDA-graph replayable.  plan (optional, [B][Hn][3] int32) receives
 * (slot, start, duration) per clip, -1 beyond the last clip. */
typedef struct {
    const float* frames;     /* [slots][max_len][per_frame] */
    const float* rewards;    /* [slots][max_len][R] */
    const int* actions;      /* [slots][max_len] */
    const int* ep_len;       /* [slots] */
    const int* n_filled;     /* device scalar */
    int slots, max_len, R;
    long long per_frame;
    int B, Hn, random_start;
    unsigned long long* rng_state;
    float* states;           /* out [B][Hn][per_frame] */
    float* rewards_out;      /* out [B][Hn][R] */
    float* dones;            /* out [B][Hn] (0 / 1) */
    long long* actions_out;  /* out [B][Hn] */
    int* plan;               /* out, optional */
} scmgan_replay_desc;
int scmgan_replay_sample(const scmgan_replay_desc* desc_host, scmgan_stream_t stream);

struct scmgan_wgrad_reduce_job;

/* Weight gradient: g[co*g_s_co + ci*g_s_ci + tap'*g_s_tap] += scale * sum_interior dy[p][co] * x[p+tap][ci],
 * tap' = flip ? 8-tap : tap.  g is fp32 and must be pre-initialised (atomically accumulated into).
 * Replaces cuDNN wgrad under loss.backward() (reference main.py:285). */
typedef struct {
    int B, H, W;
    const void* dy; /* gradient plane (only its interior is read) */
    int dy_cs, dy_c_off, cout;
    const void* x; /* input plane (halo included) */
    int x_cs, x_c_off, cin;
    float* g;
    long long g_s_co, g_s_ci, g_s_tap;
    int flip;
    int co_valid, ci_valid;
    float scale;
    void* workspace; /* optional split-K scratch (scmgan_wgrad_workspace_bytes()); with it the reduction is a second,
                        deterministic kernel instead of fp32 atomics */
    long long workspace_bytes;
    float* db; /* optional bias gradient: db[co] += sum over interior pixels of dy[p][co]; written for
                  co < co_valid rounded up to a multiple of 8 (the buffer must hold that many floats) */
    /* Deferred split-K reduction (optional).  With defer_jobs != NULL the call launches only the main kernel(s) and
     * appends the reduction(s) they need to defer_jobs[*defer_count ...] (host array of defer_cap entries); the caller
     * runs them later - e.g. on a second stream, overlapping the next layer - with scmgan_wgrad_reduce().  Partials of
     * different deferred launches must not share scratch: the call sub-allocates `workspace` from byte offset
     * *workspace_cursor (in/out, host). */
    struct scmgan_wgrad_reduce_job* defer_jobs;
    int defer_cap;
    int* defer_count;
    long long* workspace_cursor;
    int dy_fmt, x_fmt; /* SCMGAN_FMT_* of the gradient plane and of the input plane */
} scmgan_wgrad_desc;

/* One split-K reduction: g[m*g_sm + n*g_sn + tap'*g_st] += scale * sum_split ws[split][tap][n][m] (m < 128), and
 * db[m] += sum_split ws_bias[split][m] when ws_bias is set.  lanes: 8 or 32 split lanes per block. */
typedef struct scmgan_wgrad_reduce_job {
    const float* ws;
    int splits, n;
    float* g;
    long long g_sm, g_sn, g_st;
    int flip, m_valid, n_valid;
    float scale;
    const float* ws_bias;
    float* db;
    int lanes;
} scmgan_wgrad_reduce_job;
int scmgan_wgrad_reduce(int count, const scmgan_wgrad_reduce_job* jobs_host, scmgan_stream_t stream);
int scmgan_conv3x3_wgrad(const scmgan_wgrad_desc* desc_host, scmgan_stream_t stream);
long long scmgan_wgrad_workspace_bytes(void);

/* S[b][c] += sum over the interior of plane channels [c_off, c_off+n); db[c] += the same summed over b.
 * Bias gradients (cuDNN bias backward) and the folded action-channel gradient. */
int scmgan_plane_colsum(const void* plane, int Cs, int c_off, int n, int B, int H, int W, float* S, float* db,
                        scmgan_stream_t stream);

/* Spectral norm: one power iteration per wrapped conv (reference spectral_normalization.py:23-35);
 * u, v updated in place, sigma written; u_save/v_save receive the post-update vectors for backward. */
typedef struct {
    const float* w;
    float* u;
    float* v;
    float* sigma;
    float* u_save;
    float* v_save;
    int rows, cols;
} scmgan_sn_layer;
int scmgan_spectral_norm_fwd(int count, const scmgan_sn_layer* layers_host, scmgan_stream_t stream);
/* `iters` successive power iterations in one launch; iteration i of a layer writes its sigma to sigma[i*sigma_stride].
 * The iteration only reads the weights, which are constant between optimiser steps, so the 5T+3 per-call iterations of
 * one training iteration (one per SpectralNorm.forward, spectral_normalization.py:66-68) can be run ahead of the
 * rollout: u, v end in the same state and call i uses the sigma the reference would have computed in it. */
int scmgan_spectral_norm_fwd_n(int count, const scmgan_sn_layer* layers_host, int iters, int sigma_stride,
                               scmgan_stream_t stream);

/* dWbar = G/sigma - (<G,Wbar>/sigma^2) u v^T  (autograd of `w / sigma.expand_as(w)`, sigma = u.(W v)); u, v as held by
 * the module at backward time (the reference's autograd graph dereferences the Parameters then).
 * accumulate != 0: out += dWbar (what autograd's AccumulateGrad does for a weight shared by the unrolled steps of
 * main.py:162-215), else out = dWbar. */
typedef struct {
    const float* g;
    const float* wbar;
    const float* u;
    const float* v;
    const float* sigma;
    float* dot; /* [1] scratch, zeroed by caller */
    float* out;
    int rows, cols;
    int accumulate;
    const float* sigma2; /* optional.  g is the gradient w.r.t. Wbar/sigma accumulated over samples whose call used
                            sigma2 (sample_scale of scmgan_conv_desc): dWbar = g/sigma - (<g,Wbar>/(sigma*sigma2)) u v^T.
                            NULL = sigma (the plain formula above). */
} scmgan_sn_bwd_layer;
int scmgan_spectral_norm_bwd(int count, const scmgan_sn_bwd_layer* layers_host, scmgan_stream_t stream);

/* Folded action channels of Transition.conv1 (reference models.py:69-73). */
int scmgan_action_bias(const float* wbar, const float* sigma, const float* bias, const float* act, int B, int Cout,
                       int L, int A, float* out, scmgan_stream_t stream);
int scmgan_action_wgrad(const float* S, const float* act, int B, int Cout, int L, int A, float* g,
                        scmgan_stream_t stream);

/* Reward regression term of the rollout loss (reference main.py:182-186):
 *   loss[0] = scale * (*scale_dev) * mean_b( mask[b] * mean_r (pred[b][r] - target[b][r])^2 ),  dpred = d loss / d pred.
 * scale_dev (device scalar or NULL = 1) carries theta = train_iter / train_iters (main.py:143,185) so that a captured
 * CUDA graph does not bake the training progress in; loss_raw (or NULL) receives the unscaled masked mean the
 * reference logs as "Rd Loss" (main.py:184).
 * target rows are target_bstride elements apart (a time slice of the [B][Hn][R] reward tensor), mask elements
 * mask_stride apart (a time slice of the [B][Hn-1] active mask).  One launch instead of six elementwise / reduction
 * kernels forward and as many backward. */
int scmgan_masked_mse(const float* pred, const float* target, long long target_bstride, const float* mask,
                      long long mask_stride, int B, int R, float scale, const float* scale_dev, float* loss,
                      float* loss_raw, float* dpred, scmgan_stream_t stream);

/* Sequence forms of the two loss kernels.  The decoder and the reward predictor are stateless, so the T decodes /
 * reward predictions of one iteration (reference main.py:181-197, once per rollout step) run as ONE batch of T*B
 * samples: x / pred are t-major ([T][B][...]), while the targets and masks stay where they are - step t of sample b is
 * found at  base + b*bstride + t*tstride  (a [B, T] window of the [B][Hn][...] input tensors).
 *   bce:  loss_t[t] += mean_b( mask[b,t] * mean_chw BCE(sigmoid(x[t,b]), y[b,t]) )           (loss_t zeroed by caller)
 *   mse:  loss[0] = scale * (*scale_dev) * sum_t mean_b( mask[b,t] * mean_r (pred-target)^2 ), loss_raw[t] unscaled */
int scmgan_bce_logits_seq(const float* x, const float* y, long long y_bstride, long long y_tstride, const float* mask,
                          long long mask_bstride, long long mask_tstride, int T, int B, long long per, float* loss_t,
                          float* dx, scmgan_stream_t stream);
int scmgan_masked_mse_seq(const float* pred, const float* target, long long target_bstride, long long target_tstride,
                          const float* mask, long long mask_bstride, long long mask_tstride, int T, int B, int R,
                          float scale, const float* scale_dev, float* loss, float* loss_raw, float* dpred,
                          scmgan_stream_t stream);

/* loss[0] += mean_b( mask[b] * mean_chw BCE(sigmoid(x), y) ); dx (optional) = d loss / d x.
 * Fuses torch.sigmoid + F.binary_cross_entropy + means (reference main.py:188-197, 310-312). */
int scmgan_bce_logits(const float* x, const float* y, long long y_bstride, const float* mask, int B, long long per,
                      float* loss, float* dx, scmgan_stream_t stream);

/* Rollout-MSE evaluation (reference measure_prediction_mse, main.py:784-836) without per-step host reads.
 * x: logits of T decoded steps, t-major [T][B][per]; step t of sample b of the frame / reward / done tensors is found at
 * base + b*bstride + t*tstride.
 *   eval_sqerr: out[t*B+b] = mean_chw (y[b,t] - sigmoid(x[t,b]))^2                                  (main.py:812-815)
 *   eval_stats: table[t] = { mean(d)*B/live, std(d)*B/live, mean(r)*B/live, std(r)*B/live, live } with
 *               mask_t = prod_{s<=t}(1-done_s), d = mask*sqerr, r = mask*(sum_r rewards - sum_r rpred)^2, live = sum mask
 *               (main.py:806-829; std is torch.std's unbiased estimate).  rpred is [T][B][R] dense. */
int scmgan_eval_sqerr(const float* x, const float* y, long long y_bstride, long long y_tstride, int T, int B,
                      long long per, float* out, scmgan_stream_t stream);
int scmgan_eval_stats(const float* sqerr, const float* rpred, const float* rewards, long long r_bstride,
                      long long r_tstride, const float* dones, long long d_bstride, long long d_tstride, int T, int B,
                      int R, float* table, scmgan_stream_t stream);

/* Decoder loss head: the decoder's last convolution, the sigmoid, F.binary_cross_entropy, the means over C,H,W and the
 * masked batch mean (reference models.py:274-287 conv2 + latent-group sum, main.py:188-197, 310-312) as ONE kernel - the
 * BCE lives in the convolution's epilogue, the logits never reach memory.
 *   conv     the last decoder layer as a 16-column head (n = 16 >= n_valid colour channels, act = NONE, no gate / add /
 *            wrap / sample head) over the T*B hidden planes of all rollout steps, t-major (conv.B = T*B);
 *            conv.out receives d loss_t / d logits as a zero-halo gradient plane (the input of scmgan_conv3x3_dgrad /
 *            scmgan_conv3x3_wgrad of the decoder backward); conv.out_f32 (optional) still receives the logits
 *   target   frames; step t of sample b at target + b*target_bstride + t*target_tstride, dense [C][H][W]
 *   mask     active mask, element (b, t) at mask + b*mask_bstride + t*mask_tstride, or NULL
 *   loss_t   [T], loss_t[t] += mean_b( mask[b,t] * mean_chw BCE(sigmoid(logits[t,b]), target[b,t]) )  (zeroed by caller)
 * scmgan_decoder_bce_bwd applies the chain rule in place: rows of step t of the gradient plane are multiplied by g[t]
 * (= d total / d loss_t, a device vector); steps with g[t] == 1 - the training step's plain sum of terms - are skipped
 * on the device without touching the plane. */
typedef struct {
    scmgan_conv_desc conv;
    const float* target;
    long long target_bstride, target_tstride;
    const float* mask;
    long long mask_bstride, mask_tstride;
    int T, B;
    float* loss_t;
    float* workspace;          /* scratch, scmgan_decoder_bce_workspace_rows() * T floats: per-warp partial sums, reduced */
    long long workspace_bytes; /* in a fixed order by a second small kernel (no atomics: bit-reproducible loss) */
} scmgan_decoder_bce_desc;
int scmgan_decoder_bce_workspace_rows(void);
int scmgan_decoder_bce_fwd(const scmgan_decoder_bce_desc* desc_host, scmgan_stream_t stream);
int scmgan_decoder_bce_bwd(void* dlogits_plane, int cs, int fmt, const float* g, int T, int B, int H, int W,
                           scmgan_stream_t stream);

/* Counterfactual regularisers (reference main.py:242-283) on fp32 latents za, zb [B][L][H*W]:
 *   mode 0, disentanglement (258-260): loss += lambda * mean_b( mask[b] * mean_l( mean_hw|za-zb| * unswapped[b][l] ) )
 *   mode 1, action control   (279-281): loss += lambda * mean_b( mask[b] * -log( mean_{l,hw}|za-zb| + 1e-3 ) )
 * rowmean [B][L] receives mean_hw|za-zb| (needed by the backward); gscale is the device scalar d(total)/d(loss). */
int scmgan_cf_loss_fwd(const float* za, const float* zb, const float* unswapped, const float* mask, int B, int L,
                       int HW, int mode, float lambda, float* rowmean, float* loss, scmgan_stream_t stream);
int scmgan_cf_loss_bwd(const float* za, const float* zb, const float* unswapped, const float* mask,
                       const float* rowmean, const float* gscale, int B, int L, int HW, int mode, float lambda,
                       float* dza, float* dzb, scmgan_stream_t stream);

/* Stand-alone Transition tail (reference models.py:103-112): p = sigmoid(x); z = (uniforms < p), or (p > 0.5) when
 * uniforms == NULL.  The training step uses the fused conv6 epilogue instead. */
int scmgan_transition_tail(const float* x, const float* uniforms, long long n, float* p, float* z,
                           scmgan_stream_t stream);

/* Reward head of RewardPredictor (reference models.py:240-250): softmax over the 3 classes of each reward,
 * p(+1) - p(-1), summed over the stride-2 valid lattice of the second conv.  y2 is that conv evaluated as a
 * stride-1 same-size conv, fp32 [B][3R][H][W]; r [B][R]; map (optional) [B][R][h2][w2].
 * The backward writes the gradient w.r.t. y2 as a bf16 plane [B][H+2][W+2][16] (zero off the lattice). */
int scmgan_reward_head_fwd(const float* y2, int B, int R, int H, int W, float* r, float* map, scmgan_stream_t stream);
int scmgan_reward_head_bwd(const float* y2, const float* dr, int B, int R, int H, int W, void* d2_plane,
                           scmgan_stream_t stream);

/* clip_grad_value_ + Adam in one multi-tensor launch (reference main.py:287-296). */
typedef struct {
    float* p;
    const float* g;
    float* m;
    float* v;
    int n;
    float clip;        /* <= 0: no clipping */
    const float* step; /* optional per-chunk device step counter (float); overrides step/step_dev */
} scmgan_adam_chunk;
int scmgan_clip_adam(int count, const scmgan_adam_chunk* chunks_host, float lr, float beta1, float beta2, float eps,
                     int step, const float* step_dev, float gscale, scmgan_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SCMGAN_H_ */
