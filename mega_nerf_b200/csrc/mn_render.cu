// mn_render_rays: the foreground inference path of render_rays (rendering.py:15-248 with bg_nerf = None, eval mode)
// as ONE C call - coarse depths -> model query -> weights -> inverse-CDF resampling -> fine query -> merge + volume
// rendering.  It only sequences the stage entry points of this library on the caller's stream, from a caller-provided
// workspace: no allocation, no host sync, ~20 kernel launches issued back to back without returning to the host
// language in between (the Python mirror spends 1.6-2.0 ms of interpreter time on the same sequence).
#include "mn_model.cuh"

namespace {

__global__ void fill_kernel(float* p, int64_t n, float v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

struct RenderPlan {
    int64_t N;
    int Sc, Sf, Sq;            // coarse samples, fine draws, samples of the fine query (Sf, or Sc + Sf under cascade)
    int out_cols;              // columns of the raw model output (rgb_dim + 1)
    size_t z_c, xyz_c, mlp_c, raw_c, w_c, z_f, z_q, xyz_f, mlp_f, raw_f, last_delta, model_ws, total;
    size_t model_ws_bytes;
};

RenderPlan make_plan(const mn_model* m, int64_t N, int Sc, int Sf, int use_cascade, int sh, int precision) {
    RenderPlan p{};
    p.N = N; p.Sc = Sc; p.Sf = Sf;
    p.Sq = Sf > 0 ? (use_cascade ? Sc + Sf : Sf) : 0;
    p.out_cols = m->nd.rgb_dim + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += mn_align(bytes); return o; };
    p.z_c = take((size_t)N * Sc * 4);
    p.xyz_c = take((size_t)N * Sc * 12);
    p.mlp_c = sh ? take((size_t)N * Sc * p.out_cols * 4) : 0;      // raw SH coefficients before the head
    p.raw_c = take((size_t)N * Sc * 16);
    p.w_c = take((size_t)N * Sc * 4);
    p.z_f = take((size_t)N * (Sf > 0 ? Sf : 1) * 4);
    p.z_q = use_cascade && Sf > 0 ? take((size_t)N * p.Sq * 4) : p.z_f;
    p.xyz_f = take((size_t)N * (p.Sq > 0 ? p.Sq : 1) * 12);
    p.mlp_f = sh ? take((size_t)N * (p.Sq > 0 ? p.Sq : 1) * p.out_cols * 4) : 0;
    p.raw_f = take((size_t)N * (p.Sq > 0 ? p.Sq : 1) * 16);
    p.last_delta = take((size_t)N * 4);
    const size_t a = mn_model_workspace_bytes(m, N * Sc, precision);
    const size_t b = p.Sq > 0 ? mn_model_workspace_bytes(m, N * p.Sq, precision) : 0;
    p.model_ws_bytes = a > b ? a : b;
    p.model_ws = take(p.model_ws_bytes);
    p.total = off + 256;
    return p;
}

}  // namespace

extern "C" {

size_t mn_render_rays_workspace_bytes(const mn_model* m, int64_t N, int coarse_samples, int fine_samples, int use_cascade,
                                      int sh_deg, int precision) {
    if (!m || N < 0 || coarse_samples < 1 || fine_samples < 0) return 0;
    return make_plan(m, N, coarse_samples, fine_samples, use_cascade, sh_deg >= 0, precision).total;
}

int mn_render_rays(mn_ctx* ctx, mn_model* m, const float* rays_d, const float* image_indices_d, int64_t N,
                   const float* z_steps_d, int coarse_samples, const float* u_fine_d, int fine_samples, int use_cascade,
                   int sh_deg, int precision, float* rgb_out_d, float* depth_out_d, float* depth_var_out_d,
                   float* rgb_coarse_out_d, void* workspace_d, size_t workspace_bytes, void* stream) {
    if (!ctx || !m || !rays_d || !z_steps_d || !rgb_out_d || N < 0 || coarse_samples < 1 || fine_samples < 0)
        return MN_ERR_INVALID;
    if (fine_samples > 0 && !u_fine_d) return mn_fail(ctx, MN_ERR_INVALID, "mn_render_rays: u_fine_d is required when fine_samples > 0");
    if (fine_samples > 0 && coarse_samples < 3) return mn_fail(ctx, MN_ERR_INVALID, "mn_render_rays: resampling needs >= 3 coarse samples");
    const mn_model_desc& d = m->d;
    const bool sh = sh_deg >= 0;
    if (sh && (d.pos_dir_dim != 0 || d.rgb_dim != 3 * (sh_deg + 1) * (sh_deg + 1)))
        return mn_fail(ctx, MN_ERR_INVALID, "mn_render_rays: sh_deg does not match the model's rgb_dim (model_utils.py:58)");
    if (!sh && d.rgb_dim != 3) return mn_fail(ctx, MN_ERR_INVALID, "mn_render_rays: rgb_dim > 3 needs sh_deg");
    if ((d.kind == 1) != (use_cascade != 0)) return mn_fail(ctx, MN_ERR_INVALID, "mn_render_rays: use_cascade must match the model kind");
    if (!use_cascade && fine_samples == 0)
        return mn_fail(ctx, MN_ERR_INVALID, "mn_render_rays: a coarse-only render composites colour only under use_cascade (rendering.py:199)");
    if (d.appearance_dim > 0 && !image_indices_d) return mn_fail(ctx, MN_ERR_INVALID, "mn_render_rays: image indices are required");
    if (N == 0) return MN_OK;
    const RenderPlan p = make_plan(m, N, coarse_samples, fine_samples, use_cascade, sh, precision);
    if (!workspace_d || workspace_bytes < p.total) return mn_fail(ctx, MN_ERR_WORKSPACE, "mn_render_rays: workspace too small");
    char* W = (char*)workspace_d;
    auto F = [&](size_t off) { return reinterpret_cast<float*>(W + off); };
    cudaStream_t st = (cudaStream_t)stream;
    const int Sc = p.Sc, Sf = p.Sf, Sq = p.Sq;
    const bool fine = Sf > 0;
    const bool want_depth = depth_out_d || depth_var_out_d;
    int rc;

    fill_kernel<<<(unsigned)mn_cdiv(N, 256), 256, 0, st>>>(F(p.last_delta), N, 1e10f);   // no background: rendering.py:33
    MN_LAUNCH_CHECK(ctx);

    // one model query on [N, S, 3] points -> raw [N, S, 4]   (rendering.py:275-334)
    auto query = [&](const float* xyz, int S, int coarse, float* mlp_out, float* raw_out) -> int {
        mn_rows rows{};
        rows.mode = 1;
        rows.x_d = xyz;
        rows.cols = 3;
        rows.dirs_d = d.pos_dir_dim > 0 ? rays_d + 3 : nullptr;
        rows.dir_stride = 8;
        rows.idx_d = d.appearance_dim > 0 ? image_indices_d : nullptr;
        rows.samples_per_ray = S;
        float* out = sh ? mlp_out : raw_out;
        int r = mn_model_forward(ctx, m, &rows, N * S, coarse, 0, nullptr, precision, out, W + p.model_ws, p.model_ws_bytes, stream);
        if (r) return r;
        if (sh) return mn_sh_to_rgb(ctx, sh_deg, mlp_out, p.out_cols, rays_d + 3, 8, S, N * S, 1, raw_out, stream);
        return MN_OK;
    };

    // ---- coarse pass (rendering.py:82-87, 190-205)
    if ((rc = mn_sample_coarse(ctx, rays_d, nullptr, z_steps_d, nullptr, 0.0f, N, Sc, F(p.z_c), F(p.xyz_c), stream))) return rc;
    if ((rc = query(F(p.xyz_c), Sc, 1, F(p.mlp_c), F(p.raw_c)))) return rc;
    if ((rc = mn_composite(ctx, F(p.raw_c), F(p.z_c), nullptr, Sc, nullptr, nullptr, nullptr, 0, F(p.last_delta), N, 0,
                           fine ? F(p.w_c) : nullptr, use_cascade ? (fine ? rgb_coarse_out_d : rgb_out_d) : nullptr,
                           (!fine && want_depth) ? (depth_out_d ? depth_out_d : F(p.w_c)) : nullptr,
                           !fine ? depth_var_out_d : nullptr, nullptr, stream)))
        return rc;
    if (!fine) return MN_OK;

    // ---- resample (rendering.py:207-223) and fine pass (:224-243)
    if ((rc = mn_sample_pdf(ctx, F(p.z_c), F(p.w_c), Sc, nullptr, u_fine_d, 0, N, Sc, Sf, F(p.z_f), nullptr, nullptr, stream))) return rc;
    if (use_cascade)
        if ((rc = mn_sort_cat(ctx, F(p.z_c), Sc, F(p.z_f), Sf, N, 0, F(p.z_q), stream))) return rc;
    if ((rc = mn_points_from_z(ctx, rays_d, F(p.z_q), N, Sq, F(p.xyz_f), stream))) return rc;
    if ((rc = query(F(p.xyz_f), Sq, 0, F(p.mlp_f), F(p.raw_f)))) return rc;
    // depth scratch when only the variance is wanted: the coarse weights are dead by now
    float* depth_dst = want_depth ? (depth_out_d ? depth_out_d : F(p.w_c)) : nullptr;
    if (use_cascade)
        return mn_composite(ctx, F(p.raw_f), F(p.z_q), nullptr, Sq, nullptr, nullptr, nullptr, 0, F(p.last_delta), N, 0, nullptr,
                            rgb_out_d, depth_dst, depth_var_out_d, nullptr, stream);
    return mn_composite(ctx, F(p.raw_f), F(p.z_f), nullptr, Sf, F(p.raw_c), F(p.z_c), nullptr, Sc, F(p.last_delta), N, 0, nullptr,
                        rgb_out_d, depth_dst, depth_var_out_d, nullptr, stream);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Fused per-ray all-gather ("fused form"): every rank stores its rays' (rgb, depth) rows straight into
// EVERY rank's result buffer through peer-mapped pointers (NVLink / NVSwitch P2P stores), instead of packing them and
// calling a collective.  The buffers are symmetric allocations exchanged by the host (torch symmetric memory in the
// Python mirror); ordering between ranks is the caller's barrier pair (see mega_nerf_b200/dist.py::PeerGather).
// ------------------------------------------------------------------------------------------------
#define MN_MAX_PEERS 16
struct PeerBufs {
    float* p[MN_MAX_PEERS];
};

__global__ void peer_gather_store_kernel(const float* __restrict__ rgb, const float* __restrict__ depth, int64_t n, int64_t row0,
                                         PeerBufs bufs, int G) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = make_float4(rgb[i * 3 + 0], rgb[i * 3 + 1], rgb[i * 3 + 2], depth ? depth[i] : 0.0f);
    for (int g = 0; g < G; ++g) reinterpret_cast<float4*>(bufs.p[g])[row0 + i] = v;   // 16-byte store per peer
}

extern "C" int mn_peer_gather_store(mn_ctx* ctx, const float* rgb_d, const float* depth_d, int64_t n, int64_t row0,
                                    const void* const* peer_bufs, int n_peers, void* stream) {
    if (!ctx || !rgb_d || !peer_bufs || n < 0 || row0 < 0 || n_peers < 1 || n_peers > MN_MAX_PEERS) return MN_ERR_INVALID;
    if (n == 0) return MN_OK;
    PeerBufs b{};
    for (int g = 0; g < n_peers; ++g) {
        if (!peer_bufs[g]) return mn_fail(ctx, MN_ERR_INVALID, "mn_peer_gather_store: null peer buffer");
        b.p[g] = (float*)peer_bufs[g];
    }
    peer_gather_store_kernel<<<(unsigned)mn_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(rgb_d, depth_d, n, row0, b, n_peers);
    MN_LAUNCH_CHECK(ctx);
    return MN_OK;
}
