/* coordx_b200 - C ABI of the B200-native CoordX split-MLP hot path.
 *
 * The reference (nexuslrf/CoordX) has no FFI layer of its own: the boundary of
 * its hot path is the Python class API of modules.py.  This header is what a
 * binding for that path would call; `coordx_b200/modules.py` is that binding
 * (ctypes), and INTEGRATION.md shows the stub a reference maintainer would add.
 * Each entry point cites the reference code it replaces (paths relative to the
 * reference checkout).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to fp32 unless its name ends in `_host`;
 *  - weights use nn.Linear layout: row-major [out, in];
 *  - feature tables are row-major [N_c, F];
 *  - the grid is the outer product of the C per-axis sample sets, flattened
 *    row-major in branch order (modules.py:248-255), P = prod N_c;
 *  - `row_begin/row_end` select a slab of axis 0 (image rows / video frames /
 *    occupancy x - the axis the reference already loops over for memory,
 *    train_3d_occupancy.py:142-144): outputs/dout cover only that slab;
 *  - nothing allocates: the caller provides `ws` of at least
 *    coordx_*_workspace_bytes(); nothing synchronises: work is enqueued on `stream`
 *    (a cudaStream_t passed as void*);
 *  - gradient buffers are ACCUMULATED into (caller zeroes them);
 *  - return value 0 = ok, otherwise a COORDX_ERR_* / cudaError code; a message is
 *    available from coordx_last_error().
 */
#ifndef COORDX_B200_H
#define COORDX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COORDX_ABI_VERSION 3
#define COORDX_MAX_BRANCHES 4
#define COORDX_MAX_LAYERS 12

/* sine and relu run on every path; the rest of the reference's table (modules.py:170-176) on the FP32 path */
enum { COORDX_ACT_SINE = 0, COORDX_ACT_RELU = 1, COORDX_ACT_SIGMOID = 2, COORDX_ACT_TANH = 3, COORDX_ACT_SELU = 4,
       COORDX_ACT_SOFTPLUS = 5, COORDX_ACT_ELU = 6 };
enum { COORDX_FUSE_SUM = 0, COORDX_FUSE_PROD = 1 };
enum { COORDX_LOSS_MSE = 0, COORDX_LOSS_BCE = 1 };
/* arithmetic of the post-fusion trunk GEMM */
enum { COORDX_PREC_FP32 = 0,   /* FFMA, 1e-5-relative parity path, every configuration   */
       COORDX_PREC_BF16 = 1,   /* tcgen05 kind::f16 bf16 operands, fp32 accumulate in TMEM */
       COORDX_PREC_FP16 = 2 }; /* same kernel, fp16 operands (3 more mantissa bits)        */
enum { COORDX_OK = 0, COORDX_ERR_ARG = -1, COORDX_ERR_UNSUPPORTED = -2, COORDX_ERR_WORKSPACE = -3,
       COORDX_ERR_CUDA = -4 };

/* Architecture + parameters of one SplitFCBlock (modules.py:137-239).
 *   pre-fusion (per branch c): coord_linears[c] (first_w/first_b, bias scale 1),
 *                              then net[0..n_pre-1] with bias / n_branches (modules.py:224-229)
 *   fusion: sum|prod over branches, [act if fusion_before_act], [* code], group-sum
 *           fuse_width -> hidden (modules.py:336-351)
 *   trunk:  net[n_pre..n_layers-1]; the last layer of net[] has no activation
 *           (outermost_linear=True, modules.py:214-215,406).                               */
typedef struct coordx_net {
  int32_t n_branches;                       /* C = split_channels                      */
  int32_t branch_in[COORDX_MAX_BRANCHES];   /* f_c after positional encoding           */
  int32_t pre_width;                        /* hidden_features * s (modules.py:187-190) */
  int32_t fuse_width;                       /* F: feature width at the fusion           */
  int32_t hidden;                           /* fusion_feat_size (modules.py:158-163)    */
  int32_t n_pre;                            /* approx_layers                            */
  int32_t n_layers;                         /* num_hidden_layers + 1 = len(net)         */
  int32_t layer_in[COORDX_MAX_LAYERS];
  int32_t layer_out[COORDX_MAX_LAYERS];
  int32_t act;                              /* COORDX_ACT_*                             */
  int32_t fusion_op;                        /* COORDX_FUSE_*                            */
  int32_t fusion_before_act;
  int32_t has_code;                         /* learn_code                               */
  float act_scale;                          /* S = 30 * act_scale (modules.py:39)       */
  const float* first_w[COORDX_MAX_BRANCHES];
  const float* first_b[COORDX_MAX_BRANCHES];
  const float* layer_w[COORDX_MAX_LAYERS];
  const float* layer_b[COORDX_MAX_LAYERS];
  const float* code;
} coordx_net;

/* Gradient destinations, same shapes as the parameters (accumulated into). */
typedef struct coordx_grads {
  float* first_w[COORDX_MAX_BRANCHES];
  float* first_b[COORDX_MAX_BRANCHES];
  float* layer_w[COORDX_MAX_LAYERS];
  float* layer_b[COORDX_MAX_LAYERS];
  float* code;
} coordx_grads;

int coordx_abi_version(void);
const char* coordx_last_error(void);
/* measurement hooks (bench.py): number of kernels this library has launched; CUDA-event timing of
 * the dominant (tcgen05 trunk) kernel alone, on the stream it was launched on. */
unsigned long long coordx_launch_count(void);
int coordx_profile_enable(int on);
int coordx_profile_last_ms(float* ms_host);
/* optional pipeline timeline of the forward trunk kernel: device buffer of 4*96*4 int64 (see profiles/), NULL = off */
int coordx_profile_set_trace(void* dev_buf);
int coordx_profile_set_bwd_trace(void* dev_buf, int pass);   /* backward kernels: pass 0 = dW, 1 = dH */
/* 1 if the tcgen05 trunk kernels cover this architecture at `precision`. */
int coordx_trunk_tc_supported(const coordx_net* net, int precision);

/* ---- per-axis branch MLP: SplitFCBlock.forward_channel, modules.py:311-327 ----------
 * coords [n, branch_in[branch]]  ->  feat [n, fuse_width]                               */
size_t coordx_branch_workspace_bytes(const coordx_net* net, int branch, int64_t n, int backward);
int coordx_branch_forward(const coordx_net* net, int branch, const float* coords, int64_t n,
                          float* feat, void* ws, size_t ws_bytes, void* stream);
/* backward of the above (what autograd does for training.py:137-138): recomputes the
 * branch activations, accumulates weight/bias grads, optionally writes dcoords [n, f_c]. */
int coordx_branch_backward(const coordx_net* net, int branch, const float* coords, int64_t n,
                           const float* dfeat, coordx_grads* grads, float* dcoords,
                           void* ws, size_t ws_bytes, void* stream);

/* training-mode pair (what autograd's saved tensors are in the reference, modules.py:26-27,39): the forward also
 * keeps every layer's pre-activation / activation in `save` (coordx_branch_workspace_bytes(net, branch, n, 1)
 * bytes, owned by the caller until the matching backward call), the backward starts from them.  The weight- and
 * bias-gradient kernels run on an internal side stream that is forked from and joined back into `stream`.        */
int coordx_branch_forward_save(const coordx_net* net, int branch, const float* coords, int64_t n,
                               float* feat, void* save, size_t save_bytes, void* stream);
int coordx_branch_backward_saved(const coordx_net* net, int branch, const float* coords, int64_t n,
                                 const float* dfeat, coordx_grads* grads, float* dcoords,
                                 void* save, size_t save_bytes, void* stream);

/* ---- fusion + trunk: SplitFCBlock.forward_fusion, modules.py:329-354 -----------------
 * feats[c] [n[c], fuse_width]  ->  out [(row_end-row_begin) * prod_{c>0} n[c], out_width]
 * The broadcast [N_0,..,N_{C-1},F] tensor of modules.py:336-341 is never materialised for
 * the whole grid: fp32 builds it in bounded chunks (whole axis-0 rows, or pieces of one
 * row when a row alone is too long) inside `ws`; the tcgen05 path builds 128-point tiles
 * in shared memory.
 * backward: 0 = forward, 1 = backward (preferred size: the tcgen05 path keeps its 16-bit
 * dZ tiles, 64 KB per 128 points, between its two passes), 2 = backward, minimal size
 * (the tcgen05 path then recomputes instead); coordx_trunk_backward takes either.       */
size_t coordx_trunk_workspace_bytes(const coordx_net* net, const int64_t* n, int64_t row_begin,
                                    int64_t row_end, int precision, int backward);
int coordx_trunk_forward(const coordx_net* net, const float* const* feats, const int64_t* n,
                         int64_t row_begin, int64_t row_end, float* out, int precision,
                         void* ws, size_t ws_bytes, void* stream);
/* The same forward with the loss of the training step fused into the last layer (training.py:133-135 calls
 * loss_fn(model_output, gt) on the [1, P, O] result: loss_functions.py:8-12 image_mse = COORDX_LOSS_MSE,
 * :56-60 occupancy_3d = COORDX_LOSS_BCE; mean reduction, inv_count = 1 / (P * O) of the grid the mean runs over).
 * gt, dout: [slab points, out_width] like out.  loss_accum[0] += loss of the slab, dout = d loss / d out.
 * On the tcgen05 path the loss is evaluated in the trunk kernel's epilogue (no extra pass over [P, O]).      */
int coordx_trunk_forward_loss(const coordx_net* net, const float* const* feats, const int64_t* n,
                              int64_t row_begin, int64_t row_end, float* out, int loss_kind, const float* gt,
                              float inv_count, float* loss_accum, float* dout, int precision,
                              void* ws, size_t ws_bytes, void* stream);
/* dout [slab points, out_width] -> trunk weight grads + dfeat[c] [n[c], fuse_width]
 * (accumulated; dfeat of axis 0 only touches rows [row_begin,row_end)).                  */
int coordx_trunk_backward(const coordx_net* net, const float* const* feats, const int64_t* n,
                          int64_t row_begin, int64_t row_end, const float* dout,
                          coordx_grads* grads, float* const* dfeat, int precision,
                          void* ws, size_t ws_bytes, void* stream);

/* ---- losses: loss_functions.py:8-12 (image_mse), :56-60 (occupancy_3d) ---------------
 * loss_accum[0] += inv_count * sum(l(out, gt));  dout = inv_count * dl/dout (nullable).  */
int coordx_loss_forward_backward(int kind, const float* out, const float* gt, int64_t numel,
                                 float inv_count, float* loss_accum, float* dout, void* stream);

/* ---- Adam (training.py:17,146; torch.optim.Adam defaults) on one flat fp32 buffer ----
 * step_host is the 1-based step count; grad_scale multiplies the gradient first
 * (clip_grad_norm_ coefficient, training.py:140-144).                                    */
int coordx_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                     int64_t numel, float lr, float beta1, float beta2, float eps,
                     int64_t step_host, const float* grad_scale_dev, void* stream);

/* ---- the whole optimiser step of training.py:140-146 for a LIST of parameter tensors, replayable from a CUDA graph:
 * optional clip_grad_norm_(max_grad_norm) (training.py:140-144; <= 0: off), torch.optim.Adam defaults (:17,146) and
 * the ExponentialLR decay (:25; lr_gamma = 1: constant).  `state` is 4 device floats owned by the caller:
 * {lr, completed steps, scratch (must start at 0), unused}; the call leaves lr *= lr_gamma and steps += 1.            */
#define COORDX_OPT_MAX_TENSORS 32
int coordx_optim_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                      float* const* exp_avg_sq, const int64_t* numel, float* state, float beta1, float beta2,
                      float eps, float max_grad_norm, float lr_gamma, void* stream);

/* ---- batch producer of the video / volume path: dataio.py:575-600 (Implicit3DWrapper.__getitem__) gathers the
 * ground truth of a training batch, data[coord_idx[:, None] * n_xy + coord_idy[None, :], :], on the CPU and copies it
 * to the GPU every iteration; here `data` [rows, width] is device-resident.
 * out[(i * m + j), :] = data[idz[i] * stride_z + idx[j], :]   (idz == NULL: one level, nz ignored).
 * idz [nz], idx [m]: int64 device arrays; the caller guarantees they are in range.                                    */
int coordx_gather_rows(const float* data, int width, const int64_t* idz, int64_t nz, int64_t stride_z,
                       const int64_t* idx, int64_t m, float* out, void* stream);

/* ---- feature-interpolated points: the gather + lerp + fusion of ray_rendering.py:133-145 (get_pts_pred, split=True)
 * idx [n_points, n_axes] fractional indices into feats[c] [n[c], feat_width]  ->  h [n_points, feat_width], the
 * fused per-point features; they go through the trunk as a 1 x n_points grid (coordx_trunk_forward with a first
 * table of one row holding the fusion's neutral element).                                                            */
int coordx_interp_fuse(int n_axes, const float* const* feats, const int64_t* n, int feat_width, int fusion_op,
                       const float* idx, int64_t n_points, float* h, void* stream);

/* ---- whole-path inference with HOST buffers (the reference's --speed_test loop,
 * train_img.py:125-135): copies the per-axis coordinates in, runs branch + trunk, copies
 * the [P, out_width] result out.  Weights stay resident on the device (model.cuda()).    */
int coordx_infer_grid_host(const coordx_net* net, const float* const* coords_host, const int64_t* n,
                           float* out_host, int precision, void* ws, size_t ws_bytes, void* stream);
size_t coordx_infer_grid_host_workspace_bytes(const coordx_net* net, const int64_t* n, int precision);

#ifdef __cplusplus
}
#endif
#endif /* COORDX_B200_H */
