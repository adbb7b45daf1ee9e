/*
 * pasnl.h -- C ABI of libpasnl_b200.so: the PointASNL point-set hot path as hand-written sm_100a CUDA.
 *
 * This is the drop-in boundary for the native launchers the reference's TF ops / Cython shim call
 * (SURVEY.md section 8b).  Every entry point:
 *   - is extern "C", takes plain pointers and sizes (DEVICE pointers unless stated), no torch types;
 *   - keeps the argument order of the reference launcher it replaces, plus a trailing cudaStream_t (as void*);
 *   - enqueues all work on that stream, never synchronises, never allocates, keeps no state between calls;
 *   - writes every element of every output (grad entry points zero-fill their outputs themselves -- the
 *     reference makes the caller cudaMemset: tf_sampling.cpp:174, tf_grouping.cpp:204, tf_interpolate.cpp:258);
 *   - returns 0 on success, PASNL_EINVAL (-22) for bad arguments, or the positive cudaError_t of the launch.
 * All tensors are dense row-major, last index fastest, float32 / int32, exactly as the reference ops take them.
 * Paths cited below are relative to the reference repository (yanx27/PointASNL).
 */
#ifndef PASNL_H_
#define PASNL_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PASNL_OK 0
#define PASNL_EINVAL (-22)

/* library identification: "pasnl-b200 <semver> sm_100a" */
const char *pasnl_version(void);
/* human-readable text for a return code of any entry point (static storage) */
const char *pasnl_error_string(int code);

/* ---------------------------------------------------------------- sampling (tf_ops/sampling) */

/* Bytes of device scratch pasnl_fps needs for this shape (0 for every cloud that fits the register-resident
 * kernels, n <= 65536).  Replaces the caller-allocated (32,n) `temp` of tf_sampling.cpp:115. */
size_t pasnl_fps_workspace_bytes(int b, int n, int m);

/* Farthest point sampling.  Replaces farthestpointsamplingLauncher(b,n,m,inp,temp,out)
 * (tf_sampling_g.cu:203-205; kernel :105-170).  xyz (b,n,3) -> out_idx (b,m).  Bit-exact contract:
 * out[.,0]=0; distance fmaf(dz,dz,fmaf(dx,dx,dy*dy)); running min; arg-max with ties broken by lowest
 * (k mod 512), then lowest k.  workspace may be NULL when pasnl_fps_workspace_bytes()==0. */
int pasnl_fps(int b, int n, int m, const float *xyz, void *workspace, int *out_idx, void *stream);

/* gather_point.  Replaces gatherpointLauncher(b,n,m,inp,idx,out) (tf_sampling_g.cu:206-208; kernel :172-181).
 * inp (b,n,3), idx (b,m) -> out (b,m,3). */
int pasnl_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out, void *stream);
/* gather_point_grad.  Replaces scatteraddpointLauncher(b,n,m,out_g,idx,inp_g) (tf_sampling_g.cu:209-211).
 * out_g (b,m,3), idx (b,m) -> inp_g (b,n,3), zero-filled here. */
int pasnl_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g, void *stream);

/* Row gather with c channels: the tf.gather_nd of sampling() (utils/pointasnl_util.py:43-49) on features.
 * src (b,n,c), idx (b,m) -> out (b,m,c); grad: out_g (b,m,c) -> src_g (b,n,c), zero-filled here. */
int pasnl_gather_rows(int b, int n, int c, int m, const float *src, const int *idx, float *out, void *stream);
int pasnl_gather_rows_grad(int b, int n, int c, int m, const float *out_g, const int *idx, float *src_g, void *stream);

/* ---------------------------------------------------------------- grouping (tf_ops/grouping, nearest_neighbors) */

/* query_ball_point.  Replaces queryBallPointLauncher(b,n,m,radius,nsample,xyz1,xyz2,idx,pts_cnt)
 * (tf_grouping_g.cu:125-128; kernel :3-36).  xyz1 (b,n,3) support, xyz2 (b,m,3) queries ->
 * idx (b,m,nsample): first nsample support indices in ascending index order with
 * max(sqrtf(d2),1e-20f) < radius, remaining slots padded with the first hit; pts_cnt (b,m) = hits (<= nsample).
 * Rows without any hit are written as 0 (the reference leaves them uninitialised). */
int pasnl_ball_query(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx,
                     int *pts_cnt, void *stream);

/* group_point / group_point_grad.  Replace groupPointLauncher(b,n,c,m,nsample,points,idx,out) and
 * groupPointGradLauncher(b,n,c,m,nsample,grad_out,idx,grad_points) (tf_grouping_g.cu:133-141; kernels :40-78).
 * points (b,n,c), idx (b,m,nsample) -> out (b,m,nsample,c); grad_points (b,n,c) zero-filled here. */
int pasnl_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out,
                      void *stream);
int pasnl_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                           float *grad_points, void *stream);

/* Exact k nearest neighbours.  Replaces cpp_knn_batch_omp(batch_data,batch_size,npts,dim,queries,nqueries,K,
 * batch_indices) (utils/nearest_neighbors/knn_.h:17-19, knn_.cxx:104-135; dim fixed to 3, indices int32 as
 * knn_query casts them, utils/pointasnl_util.py:30) and selectionSortLauncher's use by knn_point
 * (tf_grouping.py:48-73).  xyz1 (b,n,3) support, xyz2 (b,m,3) queries -> idx (b,m,k) ascending by
 * (squared distance ((dx*dx+dy*dy)+dz*dz) evaluated without FMA, index); dist (b,m,k) squared distances or NULL.
 * Requires 1 <= k <= min(n, 128). */
int pasnl_knn(int b, int n, int m, int k, const float *xyz1, const float *xyz2, int *idx, float *dist, void *stream);

/* Fused neighbourhood gather of the SA-layer front end (utils/pointasnl_util.py:63-74 grouping() followed by
 * :248-249): out[b,j,t,:] = [ xyz[b,i,:] - center[b,j,:] | xyz[b,i,:] | feat[b,i,:] ], i = idx[b,j,t].
 * xyz (b,n,3); feat (b,n,c) or NULL with c=0; idx: m rows of nsample indices, consecutive rows idx_stride ints
 * apart (>= nsample; lets a caller pass the first-k view of a wider index tensor); center (b,m,3) or NULL
 * (then the leading 3 "relative" channels are omitted).  out (b,m,nsample,(center?3:0)+3+c). */
int pasnl_group_concat(int b, int n, int c, int m, int nsample, const float *xyz, const float *feat, const int *idx,
                       int idx_stride, const float *center, float *out, void *stream);
/* Backward of pasnl_group_concat: grad_out (b,m,nsample,(has_center?3:0)+3+c) ->
 * grad_xyz (b,n,3) (relative + absolute parts), grad_feat (b,n,c) (or NULL), grad_center (b,m,3) (or NULL,
 * required iff has_center) = -sum_t grad_rel.  All outputs zero-filled here. */
int pasnl_group_concat_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                            int idx_stride, int has_center, float *grad_xyz, float *grad_feat, float *grad_center,
                            void *stream);

/* Max over the neighbour axis: out[g,c] = max_k x[g,k,c] (tf.reduce_max(new_point, axis=[2]), utils/pointasnl_util.py:258;
 * also the group_all max pooling of utils/pointnet_util.py:125).  x (groups,k,c) -> out (groups,c).  The gradient is
 * split evenly among tied maxima, as TF's reduce_max gradient does. */
int pasnl_maxpool_k(long groups, int k, int c, const float *x, float *out, void *stream);
int pasnl_maxpool_k_grad(long groups, int k, int c, const float *x, const float *out, const float *grad_out,
                         float *grad_x, void *stream);

/* PointConv-style contraction of the SA / decoding layers (tf.transpose + tf.matmul, utils/pointasnl_util.py:272-273,
 * 330-333): out[g,ch,w] = sum_t x[g,t,ch] * w[g,t,w].  x (groups,k,c), w (groups,k,32) -> out (groups,c,32).
 * k % 4 == 0, c % 4 == 0.  Backward: grad_out (groups,c,32) -> grad_x (groups,k,c), grad_w (groups,k,32). */
int pasnl_pointconv_fwd(int groups, int k, int c, const float *x, const float *w, float *out, void *stream);
int pasnl_pointconv_bwd(int groups, int k, int c, const float *x, const float *w, const float *grad_out, float *grad_x,
                        float *grad_w, void *stream);
/* The same with an operand prologue: x holds the PRE-batch-norm output of the producing layer and is replaced on chip by
 * act(x * x_scale[ch] + x_shift[ch]) (x_act 0 none / 1 relu / 2 leaky_relu(0.2); c floats each, 16-byte aligned; NULL =
 * no prologue).  grad_x is the gradient with respect to the TRANSFORMED x (feed it to pasnl_bn_act_bwd). */
int pasnl_pointconv_fwd_bn(int groups, int k, int c, const float *x, const float *w, float *out, const float *x_scale,
                           const float *x_shift, int x_act, void *stream);
int pasnl_pointconv_bwd_bn(int groups, int k, int c, const float *x, const float *w, const float *grad_out, float *grad_x,
                           float *grad_w, const float *x_scale, const float *x_shift, int x_act, void *stream);

/* ---------------------------------------------------------------- interpolation (tf_ops/3d_interpolation) */

/* three_nn.  Replaces threenn_cpu(b,n,m,xyz1,xyz2,dist,idx) (tf_interpolate.cpp:60-103).
 * xyz1 (b,n,3) unknown, xyz2 (b,m,3) known -> dist (b,n,3) SQUARED distances ascending, idx (b,n,3);
 * strict '<' cascade: the earliest index wins ties; missing neighbours (m<3): dist=+inf, idx=0. */
int pasnl_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx, void *stream);
/* three_interpolate.  Replaces threeinterpolate_cpu(b,m,c,n,points,idx,weight,out) (tf_interpolate.cpp:107-127).
 * points (b,m,c), idx (b,n,3), weight (b,n,3) -> out (b,n,c) = (p1*w1 + p2*w2) + p3*w3, no FMA (bit-exact). */
int pasnl_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx, const float *weight,
                            float *out, void *stream);
/* three_interpolate_grad.  Replaces threeinterpolate_grad_cpu(b,n,c,m,grad_out,idx,weight,grad_points)
 * (tf_interpolate.cpp:131-153).  grad_points (b,m,c) zero-filled here. */
int pasnl_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx,
                                 const float *weight, float *grad_points, void *stream);

/* ---------------------------------------------------------------- Adaptive Sampling (utils/pointasnl_util.py:112-173) */

/* SampleWeights attention core (pointasnl_util.py:137-146).  For each of `groups` (= B*S) neighbourhoods with k
 * members and bottleneck width d:  W = softmax_rows(Q Kf^T * scale)  (k x k),  out = W V.
 * q (groups,k,d); kv (groups,k,2d) -- the conv_kv_ds output unsliced, Kf = kv[...,:d], V = kv[...,d:];
 * out (groups,k,d); attn (groups,k,k) = W, kept for the backward.  1 <= k <= 32, 1 <= d <= 256. */
int pasnl_as_core_fwd(int groups, int k, int d, const float *q, const float *kv, float scale, float *out, float *attn,
                      void *stream);
int pasnl_as_core_bwd(int groups, int k, int d, const float *q, const float *kv, const float *attn, float scale,
                      const float *grad_out, float *grad_q, float *grad_kv, void *stream);

/* AdaptiveSampling re-sampling (pointasnl_util.py:155, 168-171).  w = softmax over the k axis of
 * logits (groups,k,1+cg);  new_xyz[g,:] = sum_t w[g,t,0] * xyz[g,t,:];  new_feat[g,ch] = sum_t w[g,t,1+ch] * feat[g,t,ch].
 * xyz / feat address member t of group g at  base + g*group_stride + t*member_stride  (strides in floats), so the
 * first-k view of a grouped tensor needs no copy.  Outputs new_xyz (groups,3), new_feat (groups,cg) and
 * w (groups,k,1+cg) (kept for the backward). */
int pasnl_as_resample_fwd(int groups, int k, int cg, const float *logits, const float *xyz, long xyz_group_stride,
                          long xyz_member_stride, const float *feat, long feat_group_stride, long feat_member_stride,
                          float *new_xyz, float *new_feat, float *w, void *stream);
/* Backward: grad_new_xyz (groups,3), grad_new_feat (groups,cg) -> grad_logits (groups,k,1+cg),
 * grad_xyz (groups,k,3), grad_feat (groups,k,cg) (dense). */
int pasnl_as_resample_bwd(int groups, int k, int cg, const float *w, const float *xyz, long xyz_group_stride,
                          long xyz_member_stride, const float *feat, long feat_group_stride, long feat_member_stride,
                          const float *grad_new_xyz, const float *grad_new_feat, float *grad_logits, float *grad_xyz,
                          float *grad_feat, void *stream);

/* ---------------------------------------------------------------- Point Non-Local cell (utils/pointasnl_util.py:175-219) */

/* Attention core of PointNonLocalCell, mode='dot' (pointasnl_util.py:195-212), flash-style: for every cloud
 * A = softmax_rows(Q Kf^T * scale) over the n support points and out = A V, without materialising A.
 * q (b,s,d); kv (b,n,2d) -- the conv_kv output unsliced (Kf | V); out (b,s,d); lse (b,s) = row log-sum-exp of the
 * scaled logits, kept for the backward.  d must be a multiple of 4, 4 <= d <= 128. */
int pasnl_nl_core_fwd(int b, int s, int n, int d, const float *q, const float *kv, float scale, float *out,
                      float *lse, void *stream);
/* Bytes of device scratch pasnl_nl_core_bwd needs. */
size_t pasnl_nl_core_bwd_workspace_bytes(int b, int s, int n, int d);
/* Backward: grad_out (b,s,d) -> grad_q (b,s,d), grad_kv (b,n,2d). */
int pasnl_nl_core_bwd(int b, int s, int n, int d, const float *q, const float *kv, const float *out, const float *lse,
                      float scale, const float *grad_out, float *grad_q, float *grad_kv, void *workspace, void *stream);

/* ---------------------------------------------------------------- dense-glue epilogue (utils/tf_util.py:120-185,512-531) */

/* Train-mode batch norm + activation over the output y (rows, c) of a 1x1 conv / fully-connected GEMM, replacing
 * tf.nn.bias_add -> tf.contrib.layers.batch_norm(eps 1e-3, biased batch variance) -> activation_fn.
 * The conv bias cancels inside batch norm, so it is NOT added to y: `bias` (c) or NULL is only used to keep
 * moving_mean equal to what TF would store (mean(y) + bias).  act: 0 none, 1 relu, 2 leaky_relu(0.2).
 * Outputs: out (rows, c); mean (c) = mean(y); var (c) = biased variance (both kept for the backward);
 * running_mean / running_var (c) or NULL are updated in place: r = decay*r + (1-decay)*batch; with
 * unbiased_moving_var != 0 the moving variance is fed var * rows/(rows-1) (TF's fused batch-norm kernel, which
 * tf.contrib.layers.batch_norm takes for rank-2 / rank-4 inputs: fully_connected and conv2d), else the biased one (conv1d).
 * workspace: pasnl_bn_workspace_bytes(c) bytes. */
size_t pasnl_bn_workspace_bytes(int c);
int pasnl_bn_act_fwd(long rows, int c, const float *y, const float *gamma, const float *beta, float eps, int act,
                     float *out, float *mean, float *var, const float *bias, float *running_mean, float *running_var,
                     float decay, int unbiased_moving_var, int stats_ready, void *workspace, void *stream);
/* stats_ready != 0: `workspace` already holds the per-channel sum (c doubles) and sum of squares (c doubles) of y,
 * accumulated by pasnl_gemm_rows_tf32x3's epilogue (col_stats): the statistics pass over y is skipped. */
/* Backward of the above: dout (rows, c) -> dy (rows, c) (the gradient w.r.t. the GEMM output), dgamma (c), dbeta (c). */
int pasnl_bn_act_bwd(long rows, int c, const float *y, const float *dout, const float *gamma, const float *beta,
                     const float *mean, const float *var, float eps, int act, float *dy, float *dgamma, float *dbeta,
                     int workspace_zeroed, void *workspace, void *stream);
/* workspace_zeroed != 0: the caller hands in a workspace that is already all-zero (e.g. a slice of one arena cleared
 * once per training step), the entry point then skips its own memset node. */

/* fp32-accurate GEMM of the 1x1-conv / fully-connected layers on tcgen05 tensor cores (3xTF32 split, fp32 TMEM
 * accumulation; replaces tf.nn.conv2d / tf.matmul of tf_util.py:120-185,318-355 and their input gradients):
 *     C[rows, n] = A[rows, k] . Bt[n, k]^T ,   A, Bt, C row-major with leading dimensions lda, ldb, ldc (floats).
 * forward: A = x, Bt = W^T;  input gradient: A = dy, Bt = W. */
int pasnl_gemm_rows_tf32x3(long rows, int k, int n, const float *A, long lda, const float *Bt, long ldb, float *C,
                           long ldc, double *col_stats, void *stream);
/* col_stats: NULL, or 2*n doubles (zeroed by the caller) into which the epilogue adds the per-column sum and sum of
 * squares of C -- the batch-norm statistics, obtained without another pass over C. */

/* The same product for the tall layers whose weights fit ONE TMEM-resident tile (k <= 128, n <= 128, rows >= 2*128*SMs,
 * A 16-byte aligned with lda % 4 == 0): activations streamed by TMA tensor copies, W^T stationary in TMEM, C^T
 * accumulated so that column statistics are thread-local (csrc/gemm_ws.cu).  pasnl_gemm_rows_tf32x3 routes eligible
 * problems here by itself; the direct entry point additionally takes W in its stored [k][n] layout (w_kn != 0: no
 * transpose pass) and an optional prologue f(a) = act(a * pro_scale[col] + pro_shift[col]) on the A operand
 * (pro_act 0 none / 1 relu / 2 leaky_relu(0.2)): the batch-norm normalise + activation of the producing layer, applied
 * on the way into the tensor core instead of in a pass of its own (utils/tf_util.py:170-185). */
int pasnl_gemm_rows_ws_eligible(long rows, int k, int n, const void *A, long lda);
/* n > 128 with k <= 128 (products bound by writing C): one 128-column weight block per CTA; same entry point. */
int pasnl_gemm_rows_ws_wide_eligible(long rows, int k, int n, const void *A, long lda);
int pasnl_gemm_rows_ws(long rows, int k, int n, const float *A, long lda, const float *W, long ldw, int w_kn, float *C,
                       long ldc, double *col_stats, const float *pro_scale, const float *pro_shift, int pro_act,
                       void *stream);

/* The same product for mid-size layers whose operands are TMA-legal (16-byte aligned, lda % 4 == ldw % 4 == 0, k >= 64,
 * n >= 64, rows >= 2048): both operands streamed by TMA tensor copies, C^T accumulated (csrc/gemm_tma.cu); contractions
 * of any length (<= 1024-wide slabs summed in fp32), statistics fused for every k.  pasnl_gemm_rows_tf32x3 routes
 * eligible problems here by itself. */
int pasnl_gemm_rows_tma_eligible(long rows, int k, int n, const void *A, long lda, const void *Wt, long ldw);
int pasnl_gemm_rows_tma(long rows, int k, int n, const float *A, long lda, const float *Wt, long ldw, float *C, long ldc,
                        double *col_stats, void *stream);

/* Weight gradient of the same layers (replaces the tf.matmul / conv2d filter gradients TF autodiff emits):
 *     dW[m, n] = sum_r X[r, m] * Y[r, n] ,   X (rows, m), Y (rows, n), dW (m, n) row-major, ldw == n.
 * X = layer input, Y = gradient w.r.t. the GEMM output.  tcgen05, 3xTF32, operands transposed on the way into shared memory,
 * row slabs reduced into dW with fp32 reductions: results are order-dependent at the 1e-7 level.  dW is zero-filled
 * here unless dw_zeroed != 0 (the caller passes memory that is already all-zero, e.g. a slice of a per-step arena). */
int pasnl_gemm_wgrad_tf32x3(long rows, int m, int n, const float *X, long ldx, const float *Y, long ldy, float *dW,
                            long ldw, int dw_zeroed, void *stream);
/* The same weight gradient with both operands streamed by TMA as MN-major tiles (SWIZZLE_128B with 32-byte atomicity: no
 * transposes, csrc/gemm_wgrad_tma.cu).  Needs 16-byte aligned operands with ldx % 4 == ldy % 4 == 0, m, n >= 32,
 * rows >= 2048; pasnl_gemm_wgrad_tf32x3 routes eligible problems here by itself. */
int pasnl_gemm_wgrad_tma_eligible(long rows, int m, int n, const void *X, long ldx, const void *Y, long ldy);
/* x_scale / x_shift (m floats each) or NULL: X is replaced by act(X * x_scale[col] + x_shift[col]) on the way into the
 * tensor core (x_act 0 none / 1 relu / 2 leaky_relu(0.2)) -- the batch norm + activation of the layer that produced X,
 * when that layer's normalised output was never materialised (see pasnl_gemm_rows_ws's prologue). */
int pasnl_gemm_wgrad_tma(long rows, int m, int n, const float *X, long ldx, const float *Y, long ldy, float *dW, long ldw,
                         int dw_zeroed, const float *x_scale, const float *x_shift, int x_act, void *stream);


/* ---------------------------------------------------------------- launch-bound glue of a training step (csrc/small_ops.cu)
 * The callers' classifier head, residual adds, channel concats, dropout, losses and optimizer as library kernels, so
 * that a step contains no cuBLAS / ATen kernel.  Reference: utils/tf_util.py:318-355,590-615;
 * utils/pointasnl_util.py:283-285; utils/pointnet_util.py:218; models/pointasnl_cls.py:55-70;
 * models/pointasnl_sem_seg_res.py:72-88; train.py:150-155 (tf.train.AdamOptimizer). */

/* out (rows <= 64, nj) += in (rows, ni) . M, with M(i, j) = W[i*ldw + j] (w_transposed == 0: y = x.W, W stored [k][n]) or
 * W[j*ldw + i] (w_transposed != 0: dx = dy.W^T with the same stored W).  `out` must be zero-initialised by the caller
 * (the split contraction is reduced with atomics). */
int pasnl_gemm_skinny(int rows, int ni, int nj, const float *in, long ldin, const float *W, long ldw, int w_transposed,
                      float *out, long ldout, void *stream);
/* dW (ni, nj) = x (rows <= 64, ni)^T . dy (rows, nj): every element written. */
int pasnl_wgrad_skinny(int rows, int ni, int nj, const float *x, long ldx, const float *dy, long ldy, float *dW, long ldw,
                       void *stream);
/* out = a + b (+ c when c != NULL), n elements (tf.add chains of the layer functions). */
int pasnl_add_n(long n, const float *a, const float *b, const float *c, float *out, void *stream);
/* out (rows, ca+cb) = [a | b] along the channel axis (tf.concat(axis=-1)); split = its gradient (ga / gb may be NULL). */
int pasnl_concat_cols(long rows, int ca, int cb, const float *a, long lda, const float *b, long ldb, float *out,
                      void *stream);
int pasnl_split_cols(long rows, int ca, int cb, const float *g, float *ga, float *gb, void *stream);
/* tf.nn.dropout: out = x / keep_prob where u(seed, *step, index) < keep_prob, else 0.  The mask is a pure function of
 * (seed, *step, index): call again with the incoming gradient as x for the backward pass.  step (device, may be NULL)
 * is the optimizer's step counter, so that replays of a captured graph draw fresh masks. */
int pasnl_dropout(long n, const float *x, float keep_prob, unsigned long long seed, const long *step, float *out,
                  void *stream);
/* Sparse soft-max cross entropy, value and gradient.  weights == NULL: loss = out_scale * mean_r ce_r (tf.reduce_mean of
 * sparse_softmax_cross_entropy_with_logits); else loss = out_scale * sum_r w_r ce_r / max(#{w_r != 0}, 1)
 * (tf.losses.sparse_softmax_cross_entropy, SUM_BY_NONZERO_WEIGHTS).  labels are int64.  fwd writes lse (rows) and the
 * workspace the backward reads; bwd: dlogits = *grad_loss * d loss / d logits. */
size_t pasnl_softmax_ce_workspace_bytes(void);
int pasnl_softmax_ce_fwd(long rows, int c, const float *logits, const long *labels, const float *weights, float out_scale,
                         float *lse, void *workspace, float *loss, void *stream);
int pasnl_softmax_ce_bwd(long rows, int c, const float *logits, const long *labels, const float *weights, float out_scale,
                         const float *lse, const void *workspace, const float *grad_loss, float *dlogits, void *stream);
/* One Adam step (TF formulation) over a flat parameter buffer with the L2 regulariser folded in: for elements with
 * l2_mask != 0 the gradient gets + l2_coeff * p and *reg_loss (double, may be NULL) += l2_coeff * p^2 / 2 (pre-update
 * values).  g is multiplied by grad_scale first (1/world after a summed all-reduce).  *step (device int64) is the number
 * of steps taken so far; it is incremented on the stream after the update, and then (loss_out != NULL)
 * *loss_out = *loss_in + *reg_loss is published and *reg_loss cleared for the next step. */
int pasnl_adam_l2(long n, float *p, const float *g, float *m, float *v, const unsigned char *l2_mask, float l2_coeff,
                  float grad_scale, float lr, float beta1, float beta2, float eps, long *step, double *reg_loss,
                  const float *loss_in, float *loss_out, void *stream);

/* Max over k <= 64 neighbours with a tie mask, and the scatter form of its gradient: grad_x[g,t,c] += grad_out[g,c] /
 * popcount(mask[g,c]) for t in mask[g,c], IN PLACE into an existing (groups,k,c) gradient (tf.reduce_max's gradient is
 * split evenly among tied maxima).  pasnl_add_cols: dst[r][0..nc) += src[r][0..nc) -- the gradient of a leading-column
 * slice (grouped_xyz = new_point[..., :3], utils/pointasnl_util.py:248-249,268) added in place. */
int pasnl_maxpool_k_mask(long groups, int k, int c, const float *x, float *out, unsigned long long *mask, void *stream);
int pasnl_maxpool_k_scatter(long groups, int k, int c, const unsigned long long *mask, const float *grad_out,
                            float *grad_x, void *stream);
int pasnl_add_cols(long rows, int nc, const float *src, float *dst, long ldd, void *stream);
int pasnl_take_cols(long rows, int nc, const float *src, long lds, float *dst, void *stream);   /* dst = src[:, :nc] */
/* dst_i (cols_i, rows_i) = src_i (rows_i, cols_i)^T for `count` matrices in one launch; table (device) holds
 * {src, dst, rows, cols} as 4 int64 per matrix; max_tiles = the largest ceil(rows/32)*ceil(cols/32). */
int pasnl_transpose_batched(int count, const long *table, int max_tiles, void *stream);

/* out[r][c] = (x ? x[r][c] : 0) + bias[c] (tf.nn.bias_add; x == NULL initialises a skinny-GEMM accumulator with the
 * bias).  pasnl_col_sum: out[c] = sum_r x[r][c] (the bias gradient). */
int pasnl_bias_add(long rows, int n, const float *x, const float *bias, float *out, void *stream);
int pasnl_col_sum(long rows, int n, const float *x, float *out, void *stream);

/* ---------------------------------------------------------------- grid subsampling
 * Replaces grid_subsampling(original_points, subsampled_points, original_features, subsampled_features,
 * original_classes, subsampled_classes, sampleDl, verbose) of utils/cpp_wrappers/cpp_subsampling/grid_subsampling/
 * grid_subsampling.cpp:5-106 (what cpp_subsampling.compute(points, features=, classes=, sampleDl=) runs,
 * ScanNet/scannet_dataset_grid.py:20-38).  points (n,3) f32, features (n,fdim) f32 or NULL, classes (n,ldim) i32 or NULL
 * -> *num_out occupied cells (device int), sub_points (<= n,3), sub_features (<= n,fdim), sub_classes (<= n,ldim): the
 * caller allocates capacity n and reads *num_out.  Barycentres / feature means are bit-exact (fp32 sums in ascending
 * point index, as the reference's sequential loop); cells are emitted in order of first appearance in the input and
 * label ties go to the smallest label (the reference leaves both to its unordered_map's iteration order). */
size_t pasnl_grid_subsample_workspace_bytes(int n);
int pasnl_grid_subsample(int n, const float *points, int fdim, const float *features, int ldim, const int *classes,
                         float sample_dl, void *workspace, int *num_out, float *sub_points, float *sub_features,
                         int *sub_classes, void *stream);

/* ---------------------------------------------------------------- ops exported by the reference's libraries but unused by its models
 * pasnl_prob_sample replaces probsampleLauncher (tf_ops/sampling/tf_sampling_g.cu:194-197; ProbSample, tf_sampling.py:13-22):
 * inp_p (b,n) non-negative weights, inp_r (b,m) uniform numbers in [0,1) -> out (b,m) = index of the first running sum
 * >= r * total; temp (b,n) receives the running sums.  The fp32 summation order is the reference's, so indices match
 * bit for bit.  pasnl_selection_sort replaces selectionSortLauncher (tf_grouping_g.cu:129-132; SelectionSort,
 * tf_grouping.py:22-32): dist (b,m,n) -> outi, out (b,m,n) whose first k entries per row are the k smallest, ascending;
 * same swap rule, so ties and the unsorted remainder land where the reference puts them. */
int pasnl_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out, void *stream);
int pasnl_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out, void *stream);

/* Batch-norm finalisation without the normalise pass: from the per-column sum / sum of squares of y (ws: 2c doubles, as
 * accumulated by a GEMM epilogue) -> mean, biased var (saved for the backward), scale = gamma / sqrt(var + eps) and
 * shift = beta - mean * scale (the operand prologue of the CONSUMING layer's kernels), and the moving-average updates of
 * pasnl_bn_act_fwd.  Used when act(bn(y)) is folded into the next layer's GEMMs instead of being written out. */
int pasnl_bn_finalize(long rows, int c, const void *ws, const float *gamma, const float *beta, float eps, const float *bias,
                      float *running_mean, float *running_var, float decay, int unbiased_moving_var, float *mean,
                      float *var, float *scale, float *shift, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PASNL_H_ */
