// Internal model representation shared by the packer, the router and the MLP kernels.
#pragma once
#include "mn_common.cuh"

#define MN_MAX_LAYERS 16
#define MN_MAX_SUB 64
#define MN_TILE 128     // rows of one tensor-core MLP tile
#define MN_BUCKET 512   // slot-space bucket alignment: four consecutive 128-row tiles (two ping-pong slots of a CTA pair) never mix sub-modules

// Offsets (in floats) of each packed tensor inside one sub-module's fp32 buffer.  All matrices are
// stored K-major ("transposed": Wt[k][n] = W[n][k]) so that consecutive output channels are contiguous.
struct PackedLayout {
    int w[MN_MAX_LAYERS], b[MN_MAX_LAYERS], kin[MN_MAX_LAYERS];
    int sigma_w, sigma_b, final_w, final_b, dira_w, dira_b, rgb_w, rgb_b, emb, aff_w, aff_b;
    int total;
};

struct NetDims {
    int layers, L, in_xyz, in_dir, app, aux, rgb_dim, rgb_in, xyz_dim, nf_xyz, nf_dir, has_dir_a, affine,
        softplus, skip_mask, app_count, app_in_dira;
};

// ------------------------------------------------------------------------------------------------
// Training tapes.  Both are tiled like the fp32 MLP kernel: tile t holds TM slots
// (TM = mn_tape_tm(L)), channel-major, i.e. value (channel c, slot r of tile t) lives at
// base[(t * C + c) * TM + r].  The ACTIVATION tape is written by the forward pass in training mode,
// the GRADIENT tape (dL/d pre-activation of every Linear) by the data-gradient kernel; the weight-
// gradient kernel contracts one against the other over the slots of each sub-module.
// ------------------------------------------------------------------------------------------------
struct TapeLayout {
    // activation tape channels
    int a_pe, a_aux, a_h, a_f, a_g, a_rgb, a_lin, a_sig, a_id, a_total;
    // gradient tape channels
    int g_z, g_final, g_dira, g_rgb, g_sig, g_total;
};

// Sub-matrices of the weights in the layout the data-gradient pass streams them: [n (reduce)][k (out)],
// i.e. the nn.Linear [out,in] storage restricted to the input columns that carry a gradient.
struct BwdLayout {
    int w[MN_MAX_LAYERS];   // [L][L]    layer i >= 1, columns of the hidden part (after the PE block on skip layers)
    int final_w;            // [L][L]
    int dira_f;             // [L/2][L]  dir_a_encoding columns 0..L-1 (the xyz_encoding_final features)
    int dira_e;             // [L/2][app] dir_a_encoding columns of the appearance embedding
    int total;
};

static inline int mn_tape_tm(int L) { return L <= 256 ? 64 : 32; }

struct mn_model {
    mn_ctx* ctx = nullptr;
    mn_model_desc d{};
    NetDims nd{};
    PackedLayout lay{};
    TapeLayout tape{};
    BwdLayout blay{};
    float* packed = nullptr;          // [n_sub * lay.total] fp32
    float* packed_bwd = nullptr;      // [n_sub * blay.total] fp32, see BwdLayout
    float* centroids_d = nullptr;     // [n_sub, 3]
    int* counters_d = nullptr;        // routing scratch: see mn_route.cu
    int max_multiplicity = 0;         // slot capacity per row for blended routing
    // tensor-core packed weights (fp16 hi / lo images), see mn_mlp_tc.cu
    void* tc_packed = nullptr;
    size_t tc_sub_bytes = 0;
    int tc_ready = 0;
    // transposed fp16 weight images of the tensor-core data-gradient chain (mn_train_tc.cuh); NULL if the shape is not covered
    void* tc_dgrad = nullptr;
    size_t tc_dgrad_sub_bytes = 0;
    int train_tc_ok = 0;
    // half-major fp16 weight images + role tables of the TMEM ping-pong kernel (mn_mlp_tp.cuh); NULL for layer_dim 512
    void* tc_tp = nullptr;
    size_t tc_tp_sub_bytes = 0;
    void* tp_prog = nullptr;
    int tp_n[4] = {0, 0, 0, 0};
};

// counters_d layout (ints)
#define CNT_COUNT 0                        // [MN_MAX_SUB]   rows routed to each sub-module
#define CNT_START (MN_MAX_SUB)             // [MN_MAX_SUB+1] tile-aligned slot offsets
#define CNT_CURSOR (2 * MN_MAX_SUB + 1)    // [MN_MAX_SUB]
#define CNT_NSLOTS (3 * MN_MAX_SUB + 1)    // [1] padded slot count (end of last bucket)
#define CNT_NPAIRS (3 * MN_MAX_SUB + 2)    // [1] routed (row, sub) pairs
#define CNT_TICKET (3 * MN_MAX_SUB + 3)    // [1] blocks of the count pass that have finished (the last one scans)
#define CNT_TOTAL (3 * MN_MAX_SUB + 4)

// Arguments common to both MLP kernels.
struct MlpArgs {
    NetDims nd;
    PackedLayout lay;
    const float* packed;      // fp32 packed weights, sub s at packed + s * lay.total
    RowSrc src;
    const int* slot_row;      // slot -> row, -1 = padding; NULL = identity
    const float* slot_w;      // slot -> blend weight; NULL = 1
    const int* counters;      // routing counters (bucket starts, n_slots) or NULL
    int n_sub;
    int fixed_sub;            // used when counters == NULL
    int64_t B;                // rows (identity mode) / slot capacity (routed mode)
    int sigma_only;
    const float* sigma_noise; // [rows] or NULL
    float* out;               // [*, out_cols]
    int out_cols;
    int scatter;              // 1: out index = row, 0: out index = slot
    float* tape;              // activation tape (training forward) or NULL
    TapeLayout tl;
};

// Arguments of the backward kernels (csrc/mn_backward.cu).
struct BwdArgs {
    NetDims nd;
    PackedLayout lay;
    BwdLayout blay;
    TapeLayout tl;
    const float* packed;       // forward weights (K-major), small heads are read from here
    const float* packed_bwd;   // see BwdLayout
    const int* slot_row;       // slot -> row, -1 = padding; NULL = identity
    const float* slot_w;       // slot -> blend weight; NULL = 1
    const int* counters;       // routing counters SAVED by the forward pass, or NULL
    int n_sub;
    int fixed_sub;
    int64_t B;                 // rows (identity mode) / slot capacity (routed mode)
    const float* grad_out;     // [rows, out_cols]
    int64_t grad_rows;         // rows of grad_out (the model call's B)
    int out_cols;
    const float* act;          // activation tape
    float* grad;               // gradient tape
    float* gw;                 // parameter gradients, [n_sub * lay.total], matrices stored [out][in] (nn.Linear layout)
};

int mn_route_build(mn_ctx* ctx, mn_model* m, const RowSrc& src, int64_t B, int64_t cap, int* slot_row, float* slot_w,
                   int* row_slots, void* scratch, cudaStream_t st);
size_t mn_route_scratch_bytes(const mn_model* m, int64_t B);   // per-row active-set masks (+ blend weights [K][B])
int mn_route_combine(mn_ctx* ctx, mn_model* m, int64_t B, const int* row_slots, const float* slot_out, int out_cols,
                     float* out, cudaStream_t st);
int mn_mlp_simt_launch(mn_ctx* ctx, const MlpArgs& a, int64_t n_tiles128, cudaStream_t st);
int mn_mlp_bwd_launch(mn_ctx* ctx, const BwdArgs& a, int64_t n_tiles128, cudaStream_t st);
int mn_mlp_tc_launch(mn_ctx* ctx, mn_model* m, const MlpArgs& a, int64_t n_tiles128, int precision, void* ws,
                     size_t ws_bytes, cudaStream_t st);
size_t mn_mlp_tc_workspace(const mn_model* m, int64_t n_tiles128, int precision);
int mn_mlp_tc_pack(mn_ctx* ctx, mn_model* m, int sub, cudaStream_t st);
int mn_mlp_tp_program(const NetDims& nd, unsigned int* table_out, int cap_entries, int* info8);   // host only (test hook)
// ---- tensor-core training path (csrc/mn_train_tc.cuh): per-tile tape records and the two passes
struct TrainTcTape {
    unsigned char* xreg;      // encoder feature tiles        [n_tiles][x_tile_bytes]
    unsigned char* act;       // activation records           [n_tiles][act_tile_bytes]
    float* f32;               // [n_tiles][5][128]: sigma pre-activation, rgb (3), image id
};
size_t mn_train_tc_x_tile_bytes(const mn_model* m);
size_t mn_train_tc_act_tile_bytes(const mn_model* m);
int mn_mlp_tc_launch_train(mn_ctx* ctx, mn_model* m, const MlpArgs& a, int64_t n_tiles128, const TrainTcTape& tape, cudaStream_t st);
size_t mn_train_tc_backward_workspace(const mn_model* m, int64_t n_tiles128);
int mn_train_tc_backward(mn_ctx* ctx, mn_model* m, const BwdArgs& a, int64_t n_tiles128, const TrainTcTape& tape, void* ws, size_t ws_bytes,
                         cudaStream_t st);
