// nn.cuh -- layer primitives (linear / conv2d forward, data-grad, weight-grad), fused Adam,
// Polyak update and the Net container that mirrors border-tch-agent's SubModels.
//
// Reference: AtariCnn border-tch-agent/src/cnn/base.rs:23-48, Mlp mlp/base.rs:13-41,
// Optimizer::backward_step opt.rs:74-83 (tch nn::Adam / AdamW = torch::optim::Adam[W]),
// track util.rs:31-45.
//
// Internal layouts (device): activations NHWC ([B*OH*OW][C] row-major matrices), conv weights
// [OC][KH][KW][IC] (c1, which reads the u8 CHW frame stack straight from the replay batch, keeps
// the reference's [OC][IC][KH][KW]), linear weights [out][in]; the first linear after the conv
// stack has its input columns permuted from the reference's (c,h,w) flat_view order to (h,w,c).
// Import/export (get_param/set_param, checkpoints, SyncModel) convert to/from the reference layout.
#pragma once
#include <functional>
#include <string>
#include <vector>
#include "common.cuh"
#include "../../include/border_b200.h"

namespace bb {

// Optional per-kernel timing (bb_agent_opt_profiled): an event after every kernel launch.
struct Profiler {
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    void clear();
};

struct Ctx {
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;
    float* ws = nullptr;  // split-K / reduction workspace
    size_t ws_floats = 0;
    unsigned int* tickets = nullptr;  // per-tile arrival counters of the in-kernel split-K finish (tma_gemm.cuh), zero between launches
    size_t n_tickets = 0;
    void alloc_scratch(size_t floats);  // ws + tickets (zeroed on `stream`)
    void free_scratch();
    // tensor-core precision of the TMA-fed GEMMs launched through this context: 3 = 3xTF32 (fp32 parity, the default),
    // 1 = one TF32 product per fp32 product (the `fast` mode, bb_agent_set_precision)
    int passes = 3;
    Profiler* prof = nullptr;
    mutable std::string phase, layer;
    void mark(const char* kernel) const;  // no-op unless prof is set
    // Independent branches of the step (target forward, weight gradients) run on side contexts:
    // own stream + own split-K workspace, ordered against this stream with events.  Null = serial.
    const Ctx* side[2] = {nullptr, nullptr};
    cudaEvent_t ev = nullptr;          // scratch event of this context
    bool concurrent() const { return side[0] && !prof; }
    void fork_to(const Ctx& s) const;   // s waits for everything enqueued on this stream so far
    void join_from(const Ctx& s) const; // this stream waits for everything enqueued on s so far
};

// ---- GEMM dispatch ---------------------------------------------------------------------------
struct GemmArgs;
enum GemmMode { G_FWD = 0, G_FWD_U8, G_NN, G_WGRAD, G_WGRAD_U8, G_WGRAD_AU8 /* tcgen05 only: u8 m-contiguous A */ };
void gemm(const Ctx& c, GemmMode mode, GemmArgs a);          // nn.cu: picks tcgen05 or CUDA-core tiles
void gemm_simt(const Ctx& c, GemmMode mode, GemmArgs a);     // nn.cu: fp32 CUDA-core tiles only
bool tc_gemm(const Ctx& c, GemmMode mode, GemmArgs& a);      // tc_gemm.cu: false => not handled
bool tma_gemm(const Ctx& c, GemmMode mode, GemmArgs& a);     // tma_gemm.cu: TMA-fed tcgen05 kernel; false => not handled
void make_lo(const Ctx& c, const float* x, float* lo, size_t n);  // lo = x - tf32_trunc(x) (the second operand plane)
GemmArgs zero_args();

// ---- primitives (all row-major fp32) -------------------------------------------------------
// The *_plane arguments are the element offsets of the tensors' lo planes (x - tf32_trunc(x), see tma_gemm.cuh);
// 0 = the tensor has none / none is wanted for the output.
// Y[M][N] = act(X[M][K] W[N][K]^T + b)
void linear_fwd(const Ctx& c, const float* X, long ldx, const float* W, const float* b, float* Y, int M, int N, int K,
                bool relu, long x_plane = 0, long w_plane = 0, long y_plane = 0);
// dX[M][K] = (dY[M][N] W[N][K]) * (mask > 0)
void linear_bwd_data(const Ctx& c, const float* dY, const float* W, float* dX, long lddx, int M, int N, int K,
                     const float* mask, long dy_plane = 0, long w_plane = 0, long dx_plane = 0);
// dW[N][K] = dY^T X ; db[N] = colsum(dY)
void linear_bwd_weight(const Ctx& c, const float* dY, const float* X, long ldx, float* dW, float* db, int M, int N,
                       int K, long dy_plane = 0, long x_plane = 0);
void colsum(const Ctx& c, const float* dY, float* db, int M, int N);

struct ConvGeom {
    int B, C, H, W, OC, KH, KW, S, OH, OW;
    bool u8_chw;         // input is the u8 [B][C][H][W] frame stack (scaled by 1/255 on load)
    const int* rowbase;  // [B*OH*OW] device
    const int* koff;     // [K] device
    // Gather-form data gradient (conv_bwd_data, tensor-core path): the transposed convolution as ONE GEMM
    // over a zero-padded copy of dY.  Rows m = (image, h/S, w/S), columns n = (h%S, w%S, c), contraction
    // k = (kh/S, kw/S, oc); every operand and the output are separable gathers.  Null => col2im path.
    const int* dg_rowbase = nullptr;  // [B*(H/S)*(W/S)] per workspace: A row offsets into the padded dY
    const int* dg_crow = nullptr;     // [B*(H/S)*(W/S)] per workspace: output row offsets into dX
    const int* dg_koff = nullptr;     // [(KH/S)*(KW/S)*OC]
    const int* dg_brow = nullptr;     // [(KH/S)*(KW/S)*OC] weight rows of the B operand
    const int* dg_bnoff = nullptr;    // [S*S*C]
    const int* dg_ccol = nullptr;     // [S*S*C]
    float* dypad = nullptr;           // [B][H/S + KH/S - 1][W/S + KW/S - 1][OC] per workspace, borders stay zero
    float* dg_wt = nullptr;           // [S*S*C][(KH/S)*(KW/S)*OC] per workspace: the weights re-laid k-contiguous per step
    int wt_ready = 0;                 // dg_wt already holds this step's weights (conv_dgrad_prepare_weights ran)
    // lo planes (element offsets; 0 = none): input X, output Y / its gradient dY, weights, input gradient dX, dypad, dg_wt
    long x_plane = 0, y_plane = 0, w_plane = 0, dx_plane = 0, dypad_plane = 0, wt_plane = 0;
    // u8 first layer only: image b of the batch is row in_ix[b] of X (the replay ring; conv1_tc.cu reads it through the list)
    const unsigned long long* in_ix = nullptr;
    int M() const { return B * OH * OW; }
    int K() const { return C * KH * KW; }
    bool dgrad_gather_ok() const {
        return !u8_chw && KH % S == 0 && KW % S == 0 && H % S == 0 && W % S == 0 && C % 4 == 0 && OC % 4 == 0;
    }
    int dg_hp() const { return H / S + KH / S - 1; }
    int dg_wp() const { return W / S + KW / S - 1; }
};
// conv1_tc.cu: dedicated tcgen05 kernel for the AtariCnn first layer; false => geometry not handled
bool conv1_fwd_tc(const Ctx& c, const ConvGeom& g, const void* X, const float* W, const float* b, float* Y, bool relu);
bool conv1_wgrad_tc(const Ctx& c, const ConvGeom& g, const float* dY, const void* X, float* dW);
bool conv1_direct_ok(const ConvGeom& g);   // both of the above take the layer: the batch need not be materialised
void conv_fwd(const Ctx& c, const ConvGeom& g, const void* X, const float* W, const float* b, float* Y, bool relu);
void conv_bwd_weight(const Ctx& c, const ConvGeom& g, const float* dY, const void* X, float* dW, float* db);
// dX[B][H][W][C] = col2im(dY W) * (mask > 0); `col` is [M][K] scratch
void conv_bwd_data(const Ctx& c, const ConvGeom& g, const float* dY, const float* W, float* col, float* dX,
                   const float* mask);
bool dgrad_gather_path(const ConvGeom& g);   // the data gradient of g runs as one gather-form GEMM (needs dg_wt)
void conv_dgrad_prepare_weights(const Ctx& c, const ConvGeom& g, const float* W);

struct Exchange;
// torch::optim::Adam / AdamW step over a flat parameter vector (one launch).
struct AdamHyper {
    double lr, beta1, beta2, eps, wd;
    bool adamw;
};
// p_lo != null: the kernel also refreshes the parameters' lo plane (operand of the TMA-fed GEMMs)
void adam_step(const Ctx& c, float* p, const float* g, float* m, float* v, size_t n, const AdamHyper& h,
               uint64_t step /* 1-based */, const float* const* peer_grads = nullptr, int world = 1, float* p_lo = nullptr,
               const Exchange* wait = nullptr, int wait_regions = 0 /* bit r: wait for region r's delivered flags */,
               const struct AdamScalars* dev_scalars = nullptr);
// The step-dependent scalars of one Adam launch (bias corrections folded with lr).  With `dev_scalars` the kernel reads them
// from device memory -- the caller writes adam_scalars(h, step) there before the launch -- so that the launch itself is
// argument-invariant and can be replayed from a CUDA graph (sac.cu).
struct AdamScalars { float bc2_sqrt, neg_step, decay, pad; };
AdamScalars adam_scalars(const AdamHyper& h, uint64_t step);
// mean of all ranks' gradients, slice-owner computes and stores it into every rank's buffer (peer memory)
void grad_reduce_scatter(const Ctx& c, const float* const* peer_grads, size_t n, int rank, int world);

// Gradient exchange of a data-parallel replica with the rendezvous FOLDED INTO the kernels (no barrier launches): the
// gradient vector is exchanged in up to two regions -- region 0 as soon as its weight gradients exist (the fully connected
// layers: 95 % of the AtariCnn, finished long before the convolution backward), region 1 at the end.  Per region ONE kernel:
// announce "my gradients are ready" in every peer's flag array, wait for all ranks, reduce this rank's 1/world slice of the
// region (pairwise-tree mean, so all ranks stay bit-identical) and store it into every rank's buffer over NVLink, then
// announce "my slice is delivered".  The optimizer kernel waits in its prologue for every rank's "delivered" flag of the
// regions it consumes.  Epochs live in device memory, so every launch is argument-invariant (CUDA-graph friendly).
struct Exchange {
    const float* grads[8];     // every rank's gradient buffer (CUDA-IPC peer pointers; [rank] = local)
    unsigned int* flags[8];    // every rank's flag array; slots xflag(kind, r), written by rank r
    unsigned int* ctr;         // local: [region] epoch delivered so far, [2 + region] block-completion counters;
                               // [4 + 2k], [5 + 2k] the same for grad_exchange_ll k; [8..27] %globaltimer stamps
    int rank, world;
    int* err;                  // sticky failure flag (common.cuh)
    long long timeout_cycles;
    size_t ll_off, ll_cap;     // flag-in-data receive area behind each gradient buffer: float offset, floats per sender slot / 2
};
__host__ __device__ inline int xflag(int kind, int r) { return 16 + kind * 8 + r; }   // kind = 2*region (+1 = delivered)
void grad_exchange(const Ctx& c, const Exchange& x, size_t lo, size_t hi, int region, int max_blocks = 0);
// Small region on the critical path (the convolution gradients, 5 % of the AtariCnn): ONE hop instead of the four of
// grad_exchange.  Every rank pushes its gradients to every peer as 16-byte {value, epoch, value, epoch} stores (each 8-byte
// half lands atomically, so a receiver that sees the epoch sees the value: no fence, no flag round trip -- the "LL" protocol),
// then reduces what it received with the same pairwise tree, so all ranks end bit-identical.  Needs hi - lo <= x.ll_cap.
// `k` (0 or 1) selects the epoch counter: two such exchanges may be in flight on different streams.
void grad_exchange_ll(const Ctx& c, const Exchange& x, size_t lo, size_t hi, int k = 0);
// dest = tau*src + (1-tau)*dest  (util.rs:43)
void track(const Ctx& c, float* dest, const float* src, size_t n, double tau, float* dest_lo = nullptr);
void fill_uniform(const Ctx& c, float* p, size_t n, float bound, uint64_t seed);
void fill_const(const Ctx& c, float* p, size_t n, float v);

// ---- Net -----------------------------------------------------------------------------------

struct ParamInfo {
    std::string name;            // tch VarStore name, e.g. "c1.weight", "mlp.ln0.bias"
    std::vector<int64_t> shape;  // reference shape
    size_t offset, numel;        // into the flat parameter vector (internal layout)
    int perm;                    // 0 none, 1 conv OIHW<->OHWI, 2 linear [out][C,H,W]<->[out][H,W,C],
                                 // 3 rows [C,H,W][cols]<->[H,W,C][cols], 4 vector [C,H,W]<->[H,W,C]
    int pc, ph, pw;              // dims for the permutation
    int fan_in;
};

struct Layer {
    int type;  // 0 linear, 1 conv
    int in_dim, out_dim;  // linear
    ConvGeom geom;        // conv (B, rowbase filled per workspace)
    bool relu;
    size_t w_off, b_off;
    size_t out_elems_per_sample;
};

// Per-(net, max batch) device buffers: activations, activation grads, col scratch, gather tables.
struct NetWorkspace {
    int max_batch = 0;
    bool with_grad = false;
    std::vector<float*> act;    // output of each layer [B][out], followed by its lo plane at + plane[i]
    std::vector<float*> dact;   // grad wrt output of each layer, same
    std::vector<long> plane;    // elements between a buffer's fp32 plane and its lo plane
    std::vector<long> dypad_plane, wt_plane;
    std::vector<int*> rowbase;  // per conv layer
    std::vector<int*> dg_rowbase, dg_crow;  // per layer (null unless the gather-form data gradient applies)
    std::vector<float*> dypad, dg_wt;
    float* col = nullptr;
    size_t col_floats = 0;
    void release();
};

class Net {
  public:
    Net() = default;
    void build(const bb_net_cfg& cfg, const std::string& prefix);  // prefix: "" or "mlp."
    size_t n_params = 0;
    int in_elems = 0, out_dim = 0;
    bool u8_input = false;
    std::vector<Layer> layers;
    std::vector<ParamInfo> params;
    std::vector<int*> koff;  // per conv layer (device), shared by all workspaces
    std::vector<int*> dg_tables;  // batch-independent tables of the gather-form data gradient (owned; see ConvGeom)

    // Custom stacks (SAC's Mlp2 actor, IQN sub-nets): append one linear layer; or two heads that
    // share an input, stored as ONE [2*out][in] layer whose halves keep their own VarStore names.
    void add_linear_layer(const std::string& name, int in, int out, bool relu);
    void add_twin_heads(const std::string& name1, const std::string& name2, int in, int out);
    void reset() { layers.clear(); params.clear(); n_params = 0; }

    void init_tables(int device);
    void alloc_workspace(NetWorkspace& w, int max_batch, bool with_grad) const;
    void init_params(const Ctx& c, float* p, uint64_t seed) const;
    // forward: input [B][in] (u8 CHW frames or float rows); returns ws.act.back()
    // p_plane != 0: the parameter vector carries a valid lo plane at p + p_plane; the layers then write the lo planes of
    // their outputs and the TMA-fed tensor-core GEMMs are used wherever both operands have one.
    // in_ix != null (u8 AtariCnn input only, see direct_input_ok): `input` is the replay ring and image b is its row in_ix[b]
    const float* forward(const Ctx& c, const float* p, const void* input, long ld_in, int B, NetWorkspace& w, long p_plane = 0,
                         const unsigned long long* in_ix = nullptr) const;
    bool direct_input_ok(int B) const;   // the first layer can read ring rows through an index list at this batch size
    // Policy::sample-sized batches (B <= 8): the whole forward as ONE cooperative kernel, a warp per output element
    // and a grid-wide barrier between layers (the per-layer GEMM launches are pure latency at B = 1).  Returns null
    // when the net / batch does not qualify (caller falls back to forward()).
    const float* forward_small(const Ctx& c, const float* p, const void* input, long ld_in, int B, NetWorkspace& w) const;
    // backward from d(output) in w.dact.back(); accumulates nothing: grads are overwritten.
    // g == nullptr skips the weight gradients (data gradient only).
    // d_input (may be null) receives the gradient wrt a float input [B][in].
    // With c.concurrent() the weight gradients of all layers but the first run on the side contexts
    // while the data-gradient chain continues on c.stream; everything is joined before returning.
    // hook(i): called once the weight gradients of layers >= i have been ENQUEUED (on c.stream and the side contexts) --
    // the data-parallel agents start exchanging those gradients there, under the rest of the backward pass
    void backward(const Ctx& c, const float* p, float* g, const void* input, long ld_in, int B, NetWorkspace& w,
                  float* d_input, long ld_din, long p_plane = 0, const unsigned long long* in_ix = nullptr,
                  const std::function<void(int)>* hook = nullptr) const;
    void free_tables();
    std::string layer_name(size_t i) const;
};

// reference layout <-> internal layout of one parameter tensor (host side)
void param_to_internal(const ParamInfo& pi, const float* ref, float* internal);
void param_to_reference(const ParamInfo& pi, const float* internal, float* ref);

}  // namespace bb
