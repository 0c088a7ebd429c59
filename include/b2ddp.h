/*
 * b2ddp.h -- C-ABI of the B200-native data-parallel training path.
 *
 * Drop-in boundary for the hot path of rensortino/DDP-Tutorial
 * (/root/reference/cifar_example_ddp.py:82-110).  The reference has no FFI of its
 * own: every entry point below replaces the torch/NCCL call that the cited
 * reference line dispatches to.  Plain C: raw device pointers, sizes, scalars and
 * cudaStream_t passed as void*.  No torch types, no C++ types, nothing throws.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; b2_last_error() returns a
 *     thread-local message for the last failing call on this thread;
 *   - all compute entries are asynchronous and ordered on the `stream` argument;
 *   - the caller owns every buffer it passes and keeps it alive until the stream
 *     work completes; the library owns what it allocates (model workspaces, comm
 *     arenas, graphs) and frees it in *_destroy / b2_comm_finalize;
 *   - one host thread per rank-process.
 */
#ifndef B2DDP_H_
#define B2DDP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_OK 0
#define B2_ERR_CUDA (-1)
#define B2_ERR_ARG (-2)
#define B2_ERR_STATE (-3)
#define B2_ERR_TIMEOUT (-4)
#define B2_ERR_UNSUPPORTED (-5)

/* model kinds: the reference's Net (cifar_example_ddp.py:23-40) and the two
 * torchvision topologies BASELINE.json substitutes at cifar_example.py:60 */
#define B2_MODEL_NET 0
#define B2_MODEL_RESNET18 1
#define B2_MODEL_RESNET50 2

/* arithmetic mode: 0 = fp32 storage + fp32 FMA (parity mode), 1 = bf16
 * activations / bf16 weight copies, fp32 accumulate, fp32 master weights */
#define B2_PREC_FP32 0
#define B2_PREC_BF16 1

typedef struct b2_model b2_model; /* opaque */

/* ---- library ------------------------------------------------------------- */
int b2_version(void);
const char *b2_last_error(void);
/* number of kernels this library has launched in this process (bench.py's
 * gpu_launches claim); monotonically increasing. */
int64_t b2_launch_count(void);
/* Diagnostic kernel trace: while active, block 0 of every kernel of this library appends (tag, %globaltimer ns)
 * records -- tag = translation unit << 24 | source line << 4 | kind (0 = kernel body begins, 1 = block 0 ends) --
 * also inside replayed CUDA graphs and on the side streams (tools/step_trace.py turns them into a timeline).
 * b2_trace_stop copies up to max_records (two uint64 each) to the host and switches the trace off. */
int b2_trace_start(int capacity);
int b2_trace_stop(unsigned long long *records, int max_records, int *n_out);

/* ---- model: replaces Net()/resnetXX() + .to(gpu) (cifar_example_ddp.py:82) - */
int b2_model_create(int kind, int batch, int num_classes, int precision, int device, b2_model **out);
int b2_model_destroy(b2_model *m);
/* flat parameter vector layout (what DDP would put in its buckets,
 * reducer.hpp:347-349): number of fp32 elements, number of tensors */
int64_t b2_model_param_count(const b2_model *m);
int b2_model_num_tensors(const b2_model *m);
/* i-th parameter tensor in torch named_parameters() order: name (static string),
 * element offset in the flat vector, logical torch shape (rank<=4) and the
 * storage strides (in elements) of that logical shape inside the flat vector --
 * conv weights are stored O,H,W,I so their O,I,H,W view is a strided view. */
int b2_model_tensor_info(const b2_model *m, int i, const char **name, int64_t *offset,
                         int *rank, int64_t shape[4], int64_t stride[4]);
/* buffers (BatchNorm running_mean / running_var, fp32; num_batches_tracked is a
 * host-side counter): same query for the flat buffer vector */
int64_t b2_model_buffer_count(const b2_model *m);
int b2_model_num_buffers(const b2_model *m);
int b2_model_buffer_info(const b2_model *m, int i, const char **name, int64_t *offset, int64_t *numel);
int64_t b2_model_steps_tracked(const b2_model *m);
/* bind caller-owned flat fp32 vectors (params, grads; buffers may be NULL) */
int b2_model_bind(b2_model *m, float *params, float *grads, float *buffers);
/* bf16 mode: refresh the bf16 compute copies of the weights from the fp32
 * masters (done automatically by b2_sgd_step_model; call after loading params) */
int b2_model_sync_weights(b2_model *m, void *stream);
/* bf16 mode: use a caller-owned bf16 mirror of the flat parameter vector (same element
 * offsets, 128-byte aligned) instead of the model's own; lets one optimizer kernel keep the
 * mirror of several models (batch sizes) current.  NULL restores the model-owned mirror. */
int b2_model_bind_bf16(b2_model *m, void *bf16_params);
/* the caller refreshed the mirror itself (b2_sgd with p_bf16) */
int b2_model_mark_weights_clean(b2_model *m);
/* tell the model that the bound fp32 parameters changed behind its back (optimizer
 * step, load_state_dict): bf16 compute copies are refreshed before the next forward */
int b2_model_mark_weights_dirty(b2_model *m);
/* net(inputs): cifar_example_ddp.py:104.  x is fp32 NCHW [batch,3,32,32] on the
 * device; logits fp32 [batch,num_classes] */
int b2_forward(b2_model *m, const float *x_nchw, float *logits, void *stream);
/* loss.backward(): cifar_example_ddp.py:106.  dlogits fp32 [batch,num_classes];
 * writes every parameter gradient into the bound flat grad vector */
int b2_backward(b2_model *m, const float *dlogits, void *stream);
/* forward/backward are replayed from CUDA graphs after two eager calls (default on); one graph
 * per distinct (input, output) pointer pair.  0 disables and drops the cached graphs. */
int b2_model_set_graphs(b2_model *m, int enable);
/* measurement: when enabled, forward/backward bracket every op with CUDA events on the
 * launching stream; collect() synchronises and returns, per category (B2_PROF_*), the summed
 * milliseconds, algorithmic FLOPs, algorithmic bytes and number of op launches since the last
 * collect.  Arrays have B2_PROF_COUNT entries. */
#define B2_PROF_CONV_FPROP 0
#define B2_PROF_CONV_DGRAD 1
#define B2_PROF_CONV_WGRAD 2
#define B2_PROF_BN_FWD 3
#define B2_PROF_BN_BWD 4
#define B2_PROF_POOL 5
#define B2_PROF_OTHER 6
#define B2_PROF_COUNT 7
int b2_model_profile(b2_model *m, int enable);
int b2_model_profile_collect(b2_model *m, double *ms, double *flops, double *bytes, int *launches);
/* roofline leg: average duration of every tcgen05 conv launch of one step (each plan launched
 * `iters` times back to back between two CUDA events on `stream`), summed per class
 * [0]=fprop [1]=dgrad [2]=wgrad, with the FLOPs those launches execute.  Clobbers gradients. */
int b2_model_time_conv_kernels(b2_model *m, int iters, double *ms3, double *flops3, int *launches3,
                               void *stream);
/* synchronises and reports B2_ERR_TIMEOUT if a fused conv+BatchNorm kernel gave up at its
 * grid barrier (its results are then invalid) */
int b2_model_check(b2_model *m);
/* debugging / parity: copy an internal activation (by op index) out as fp32 NCHW */
int b2_model_num_ops(const b2_model *m);
int b2_model_read_activation(b2_model *m, int op, int want_grad, float *out, int64_t capacity,
                             int64_t *numel, int shape[4], void *stream);

/* kernel-level parity hook: ONE convolution pass on caller buffers (NHWC bf16 activations,
 * fp32 O,H,W,I master weights) through the tcgen05 path (use_umma=1) or the CUDA-core path
 * (use_umma=0).  mode 0: out = y bf16 [N,Ho,Wo,Co]; 1: out = dx bf16 [N,Hi,Wi,Ci] (acc: +=);
 * 2: out = dw fp32 [Co][k*k*Ci].  Synchronises the stream. */
int b2_conv_test(int N, int Hi, int Wi, int Ci, int Co, int k, int stride, int pad, int mode,
                 int use_umma, int acc, const void *x_bf16, const float *w_master,
                 const void *dy_bf16, void *out, void *stream);
/* one fprop launch with the fused BatchNorm(+residual, ReLU) epilogue, exactly as the model issues it for
 * every conv+BN pair (torchvision resnet.py:92-94: conv -> bn -> relu): z = conv(x), y = relu?(bn_train(z) + res).
 * info3 (optional) returns {output-channel tile, persistent grid, tile count}; timeline13 (optional, host) returns
 * clock64() of CTA 0 at 13 points of the kernel (setup, first data, last MMA, epilogue phases, grid barrier, ...). */
int b2_conv_bn_test(int N, int Hi, int Wi, int Ci, int Co, int k, int stride, int pad, int relu,
                    const void *x_bf16, const float *w_master, const void *res_bf16, const float *gamma,
                    const float *beta, void *z_out, void *y_out, float *mean_out, float *invstd_out,
                    int *info3, long long *timeline13, void *stream);

/* ---- next rows: input transform and eval counters -----------------------------
 * transforms.ToTensor() + Normalize(mean, std) (cifar_example_ddp.py:61-63) on the device:
 * uint8 [N][H][W][3] (HWC, host values 0..255) -> fp32 NCHW.  mean3/std3 are HOST pointers. */
int b2_preprocess_u8(const void *x_u8_nhwc, float *y_nchw, int N, int H, int W, int C,
                     const float *mean3, const float *std3, void *stream);
/* eval loop body (cifar_example_ddp.py:131-133): counts[0] += correct top-1, counts[1] += batch;
 * counts is a device int64[2] the caller zeroes once. */
int b2_argmax_count(const float *logits, const int64_t *labels, int batch, int classes,
                    long long *counts, void *stream);

/* ---- criterion: nn.CrossEntropyLoss() fwd+bwd (cifar_example_ddp.py:85,105) -- */
int b2_ce_loss_grad(const float *logits, const int64_t *labels, int batch, int classes,
                    float *loss_out, float *dlogits, void *stream);
/* same, and additionally running2[0] += loss, running2[1] += 1 on the device: the reference's
 * `running_loss += loss.item()` (cifar_example_ddp.py:110) without a host sync every step; the
 * caller reads running2 back every print interval (:111-114).  running2 may be NULL. */
int b2_ce_loss_grad_running(const float *logits, const int64_t *labels, int batch, int classes,
                            float *loss_out, float *dlogits, float *running2, void *stream);

/* ---- optimizer: optim.SGD(lr, momentum).step() (cifar_example_ddp.py:86,107) -- */
/* p,g,m fp32 [n]; first_step!=0 => m = g (sgd.py:_single_tensor_sgd clones the
 * grad on the first step), else m = momentum*m + g; p -= lr*m.  grad_scale
 * multiplies g first (1/world or loss-scale inverse; pass 1.0f for none). */
int b2_sgd(float *p, const float *g, float *m, int64_t n, float lr, float momentum,
           float grad_scale, int first_step, void *p_bf16 /* optional bf16 mirror of p */, void *stream);

/* ---- peer-memory communicator: replaces init_process_group('nccl') + DDP's
 * Reducer all-reduce (cifar_example_ddp.py:42-58,83; reducer.hpp:524) ---------- */
#define B2_IPC_HANDLE_BYTES 64
int b2_comm_init(int rank, int world, int device);
int b2_comm_finalize(void);
/* allocate `bytes` of peer-mappable device memory (library-owned) */
int b2_comm_alloc(int64_t bytes, void **ptr);
int b2_comm_free(void *ptr);
/* export an IPC handle for a b2_comm_alloc'd pointer; import a peer's handle */
int b2_comm_export(void *ptr, void *handle_out /* B2_IPC_HANDLE_BYTES */);
int b2_comm_import(const void *handle, void **peer_ptr);
int b2_comm_unimport(void *peer_ptr);
/* register the per-rank buffers of one "channel": grads[w], params[w] and
 * flags[w] (uint32 signal pads, >= B2_FLAG_WORDS words, zero-initialised) are the
 * world-size arrays of mapped pointers (entry `rank` is the local one). */
#define B2_FLAG_WORDS 2048
int b2_comm_register(void *const *grads, void *const *params, void *const *flags, int64_t n_elems);
/* dist.barrier() (cifar_example_ddp.py:58): flag barrier in peer memory */
int b2_comm_barrier(void *stream);
/* size in bytes of one rank's flag pad (== B2_FLAG_WORDS * 4) */
/* ---- NVLS (in-switch reduction) support: VMM allocations + multicast objects --------------------------------
 * Replaces what NCCL sets up for the NVLS all-reduce behind cifar_example_ddp.py:57,83,106.  Host protocol
 * (parallel.py): every rank b2_vmm_alloc()s its gradient / parameter vectors, sends the file descriptors to its
 * peers over Unix sockets, b2_vmm_import()s theirs (plain peer pointers); rank 0 b2_mc_create()s one multicast
 * object per vector and distributes its descriptor; every rank b2_mc_import()s (rank 0 keeps its handle),
 * b2_mc_add_device()s, and -- once ALL ranks have added -- b2_mc_bind_and_map()s; the resulting multicast
 * addresses go to b2_comm_register_multicast().  Falls back to the peer load/store kernels when
 * b2_vmm_query() reports no multicast support. */
int b2_vmm_query(int device, int *multicast_supported, int64_t *granularity);
int b2_vmm_alloc(int device, int64_t bytes, void **ptr, int64_t *handle_out, int *fd_out);
int b2_vmm_import(int device, int fd, int64_t bytes, void **peer_ptr);
int b2_mc_create(int world, int64_t bytes, int64_t *mc_handle, int *fd_out);
int b2_mc_import(int fd, int64_t *mc_handle);
int b2_mc_add_device(int64_t mc_handle, int device);
int b2_mc_bind_and_map(int64_t mc_handle, int64_t mem_handle, int64_t bytes, int device, void **mc_ptr);
int b2_vmm_close_fd(int fd);
int b2_vmm_release_all(void);
int b2_comm_register_multicast(void *mc_grads, void *mc_params);
int b2_comm_set_nvls(int enable);
int b2_comm_nvls_active(void);
int b2_comm_flag_bytes(void);
/* synchronises the device and reports B2_ERR_TIMEOUT if any comm kernel gave up waiting
 * for a peer flag (a rank died or the ranks' launch sequences diverged) */
int b2_comm_check(void);
/* plain all-reduce(sum) * scale of the registered grad buffer, in place on every
 * rank (for the bandwidth sweep, BASELINE config 5).  algo: 0 auto, 1 one-shot,
 * 2 two-shot */
int b2_allreduce(int64_t n_elems, float scale, int algo, void *stream);
/* same on the sub-range [offset, offset+n_elems) (offset % 4 == 0): one DDP bucket */
int b2_allreduce_range(int64_t offset, int64_t n_elems, float scale, int algo, void *stream);
/* fused: all-reduce(sum)/world of the registered grads + SGD-momentum update of
 * the registered params (m is local, fp32 [n]): Reducer + optimizer.step() in one
 * kernel (cifar_example_ddp.py:106-107). */
int b2_allreduce_sgd(float *m, int64_t n_elems, float lr, float momentum, int first_step,
                     int algo, void *p_bf16 /* optional local bf16 mirror */, void *stream);
int b2_allreduce_sgd_range(float *m, int64_t offset, int64_t n_elems, float lr, float momentum,
                           int first_step, int algo, void *p_bf16, void *stream);
/* DDP inside the model: with a communicator registered, b2_backward itself launches the fused
 * all-reduce + SGD of every gradient bucket on a side stream as soon as the bucket's last
 * gradient has been produced (overlap with the rest of backward, reducer.hpp:524 semantics) and
 * joins before returning to the caller's stream; optimizer.step() then has nothing left to
 * launch.  momentum: fp32 [param_count].  enable=0 restores the plain backward. */
int b2_model_set_ddp(b2_model *m, int enable, float *momentum_buf, float lr, float momentum, int algo,
                     int64_t bucket_elems);
/* single process: same bucketing, the per-bucket kernel is the plain SGD-momentum update (no exchange):
 * the optimizer pass of the first (largest) buckets runs under the rest of backward. */
int b2_model_set_fused_sgd(b2_model *m, int enable, float *momentum_buf, float lr, float momentum,
                           int64_t bucket_elems);
/* step 0 clones the gradient into the momentum buffer (sgd.py): tell the model which step it is */
int b2_model_ddp_first_step(b2_model *m, int first);
int b2_model_num_buckets(const b2_model *m);
/* the gradient-bucket schedule for a given bucket size, without enabling anything (returns the
 * number of buckets; arrays receive up to max_buckets entries: element offset, element count, index
 * of the op whose backward completes the bucket).  Works on metadata-only models. */
int b2_model_bucket_plan(b2_model *m, int64_t bucket_elems, int max_buckets, int64_t *offsets,
                         int64_t *counts, int *after_op);
/* b2_model_create with device < 0 builds a metadata-only instance (no CUDA call): parameter table,
 * op list and schedules can be queried on a host without a GPU; forward/backward return an error. */
int64_t b2_model_workspace_bytes(const b2_model *m);
/* op i of the static graph: type (0 input, 1 conv/linear, 2 batchnorm, 3 maxpool, 4 avgpool),
 * geom = {N,Hi,Wi,Ci,Ho,Wo,Co,kh,kw,stride,pad}, flags = {tcgen05 conv, tcgen05 stem, fused BN, fc head} */
int b2_model_op_info(const b2_model *m, int op, int *type, int geom[11], int flags[4]);

/* ---- CUDA-graph capture of a step (launch-latency removal) ---------------- */
int b2_graph_begin(void *stream);
int b2_graph_end(void *stream, void **graph_exec);
int b2_graph_launch(void *graph_exec, void *stream);
int b2_graph_destroy(void *graph_exec);

#ifdef __cplusplus
}
#endif
#endif /* B2DDP_H_ */
