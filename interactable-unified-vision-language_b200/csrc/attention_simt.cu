// fp32-math attention with decomposed relative-position bias (validation mode; also the bring-up path for bf16).
//
// Restates Attention.forward + add_decomposed_rel_pos (image_encoder.py:239-255, 340-376) fused with
// window_partition / window_unpartition (image_encoder.py:258-304):
//   * the qkv tensor stays in token order [B*g*g, 3D]; a CTA gathers the ws x ws window it works on by index
//     arithmetic, so neither partition nor unpartition materialises;
//   * pad positions of edge windows (token y or x >= g) are REAL keys with k = b_k, v = b_v (pad tokens are zero
//     AFTER norm1, image_encoder.py:183-187,274) and are included in the softmax; pad QUERIES are skipped because
//     window_unpartition crops them (image_encoder.py:302-303);
//   * bias(q,k) = q . rel_h[qh-kh+ws-1] + q . rel_w[qw-kw+ws-1] with the UNSCALED q (image_encoder.py:249), built
//     per query tile as two [32][ws] tables in shared memory — the (S,S) bias is never materialised;
//   * online softmax in fp32, exact expf.
// One CTA = 32 queries of one (image, window, head); 4 threads per query row; key tiles of 64.
#include "common.cuh"

namespace svb {
namespace {

constexpr int QT = 32, KT = 64, MAXHD = 128, NT = 128;

template <typename T>
__global__ void __launch_bounds__(NT)
attention_simt_kernel(AttnParams p) {
    extern __shared__ float sm[];
    // token grid gh x gw, windows wsh x wsw (the whole grid for global attention)
    const int hd = p.hd, D = p.heads * p.hd;
    const int gh = p.grid_h > 0 ? p.grid_h : p.grid, gw = p.grid_h > 0 ? p.grid_w : p.grid;
    const int wsh = p.grid_h > 0 ? p.ws_h : p.ws, wsw = p.grid_h > 0 ? p.ws_w : p.ws;
    const int hdp = hd + 1;
    float* Qs = sm;                          // [QT][hdp]
    float* Ks = Qs + QT * hdp;               // [KT][hdp]
    float* Vs = Ks + KT * hdp;               // [KT][hd]
    float* Ps = Vs + KT * hd;                // [QT][KT+1]
    float* Bh = Ps + QT * (KT + 1);          // [QT][wsh]
    float* Bw = Bh + QT * wsh;               // [QT][wsw]

    const int tid = threadIdx.x;
    const int r = tid >> 2, g4 = tid & 3;
    const int head = blockIdx.y;
    const int nwy = (gh + wsh - 1) / wsh, nwx = (gw + wsw - 1) / wsw;
    const int nwin = nwy * nwx;
    const int b = blockIdx.z / nwin, win = blockIdx.z % nwin;
    const int wy = win / nwx, wx = win % nwx;
    const int S = wsh * wsw;
    const size_t ntok = (size_t)gh * gw;
    const int q0 = blockIdx.x * QT;
    const T* qkv = reinterpret_cast<const T*>(p.qkv);
    const size_t ld = (size_t)3 * D;
    const float scale = rsqrtf((float)hd);

    // ---- Q tile (unscaled) ----
    for (int i = tid; i < QT * hd; i += NT) {
        const int rr = i / hd, c = i % hd;
        const int qi = q0 + rr;
        float v = 0.f;
        if (qi < S) {
            const int y = wy * wsh + qi / wsw, x = wx * wsw + qi % wsw;
            if (y < gh && x < gw) v = to_float(qkv[((size_t)b * ntok + (size_t)y * gw + x) * ld + head * hd + c]);
        }
        Qs[rr * hdp + c] = v;
    }
    __syncthreads();
    // ---- decomposed rel-pos tables for this query tile ----
    for (int i = tid; i < QT * (wsh + wsw); i += NT) {
        const int rr = i / (wsh + wsw), j = i % (wsh + wsw);
        const int qi = q0 + rr;
        const bool is_w = j >= wsh;
        const int kk = is_w ? j - wsh : j;
        float acc = 0.f;
        if (qi < S) {
            const int qpos = is_w ? (qi % wsw) : (qi / wsw);
            const float* tab = (is_w ? p.rel_w : p.rel_h) + (size_t)(qpos - kk + (is_w ? wsw : wsh) - 1) * hd;
            for (int c = 0; c < hd; ++c) acc = fmaf(Qs[rr * hdp + c], __ldg(tab + c), acc);
        }
        if (is_w) Bw[rr * wsw + kk] = acc;
        else Bh[rr * wsh + kk] = acc;
    }

    float m_run = -INFINITY, l_run = 0.f;
    float o[MAXHD / 4];
#pragma unroll
    for (int t = 0; t < MAXHD / 4; ++t) o[t] = 0.f;
    const int qi = q0 + r;
    const int qh = qi / wsw, qw = qi % wsw;

    for (int k0 = 0; k0 < S; k0 += KT) {
        __syncthreads();    // previous tile fully consumed (also orders the Bh/Bw writes on the first pass)
        for (int i = tid; i < KT * hd; i += NT) {
            const int kr = i / hd, c = i % hd;
            const int ki = k0 + kr;
            float kv = 0.f, vv = 0.f;
            if (ki < S) {
                const int y = wy * wsh + ki / wsw, x = wx * wsw + ki % wsw;
                if (y < gh && x < gw) {
                    const T* base = qkv + ((size_t)b * ntok + (size_t)y * gw + x) * ld + head * hd + c;
                    kv = to_float(base[D]);
                    vv = to_float(base[2 * D]);
                } else {   // pad token: qkv(0) = bias; round through T like a stored activation would be
                    kv = to_float(from_float<T>(__ldg(p.qkv_bias + D + head * hd + c)));
                    vv = to_float(from_float<T>(__ldg(p.qkv_bias + 2 * D + head * hd + c)));
                }
            }
            Ks[kr * hdp + c] = kv;
            Vs[kr * hd + c] = vv;
        }
        __syncthreads();
        // ---- scores for keys g4, g4+4, ... ----
        float s[KT / 4];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < KT / 4; ++j) {
            const int kr = g4 + 4 * j;
            const int ki = k0 + kr;
            float acc = 0.f;
            for (int c = 0; c < hd; ++c) acc = fmaf(Qs[r * hdp + c], Ks[kr * hdp + c], acc);
            if (ki < S && qi < S) {
                acc = acc * scale + Bh[r * wsh + ki / wsw] + Bw[r * wsw + ki % wsw];
            } else {
                acc = -INFINITY;
            }
            s[j] = acc;
            mx = fmaxf(mx, acc);
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float m_new = fmaxf(m_run, mx);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = (m_run == -INFINITY) ? 0.f : expf(m_run - m_use);
        float psum = 0.f;
#pragma unroll
        for (int j = 0; j < KT / 4; ++j) {
            const float pj = (s[j] == -INFINITY) ? 0.f : expf(s[j] - m_use);
            Ps[r * (KT + 1) + g4 + 4 * j] = pj;
            psum += pj;
        }
        psum += __shfl_xor_sync(0xffffffffu, psum, 1);
        psum += __shfl_xor_sync(0xffffffffu, psum, 2);
        l_run = l_run * alpha + psum;
        m_run = m_new;
        __syncwarp();
        // ---- O += P V for columns g4, g4+4, ... ----
#pragma unroll
        for (int t = 0; t < MAXHD / 4; ++t) {
            const int c = g4 + 4 * t;
            if (c < hd) {
                float acc = o[t] * alpha;
                for (int kr = 0; kr < KT; ++kr) acc = fmaf(Ps[r * (KT + 1) + kr], Vs[kr * hd + c], acc);
                o[t] = acc;
            }
        }
    }
    if (qi < S) {
        const int y = wy * wsh + qh, x = wx * wsw + qw;
        if (y < gh && x < gw) {
            T* out = reinterpret_cast<T*>(p.out) + ((size_t)b * ntok + (size_t)y * gw + x) * D + head * hd;
            const float inv = 1.f / l_run;
#pragma unroll
            for (int t = 0; t < MAXHD / 4; ++t) {
                const int c = g4 + 4 * t;
                if (c < hd) out[c] = from_float<T>(o[t] * inv);
            }
        }
    }
}

}  // namespace

int attention_simt(const AttnParams& p, bool is_bf16, cudaStream_t stream) {
    SVB_REQUIRE(p.hd <= MAXHD && p.hd > 0, "attention_simt: head_dim %d not supported (max %d)", p.hd, MAXHD);
    const int gh = p.grid_h > 0 ? p.grid_h : p.grid, gw = p.grid_h > 0 ? p.grid_w : p.grid;
    const int wsh = p.grid_h > 0 ? p.ws_h : p.ws, wsw = p.grid_h > 0 ? p.ws_w : p.ws;
    SVB_REQUIRE(gh > 0 && gw > 0 && wsh > 0 && wsw > 0 && wsh <= gh && wsw <= gw, "attention_simt: bad window %d x %d for grid %d x %d", wsh,
                wsw, gh, gw);
    const int nwin = ((gh + wsh - 1) / wsh) * ((gw + wsw - 1) / wsw);
    const int S = wsh * wsw;
    const int hdp = p.hd + 1;
    const size_t smem = sizeof(float) * ((size_t)QT * hdp + (size_t)KT * hdp + (size_t)KT * p.hd + (size_t)QT * (KT + 1) +
                                         (size_t)QT * (wsh + wsw));
    SVB_REQUIRE(smem <= 227 * 1024, "attention_simt: window %d x %d needs %zu bytes of shared memory", wsh, wsw, smem);
    dim3 grid((S + QT - 1) / QT, p.heads, p.batch * nwin);
    const double D_ = (double)p.heads * p.hd;
    ProfScope prof((wsh == gh && wsw == gw) ? PC_ATTN_GLOBAL : PC_ATTN_WIN,
                   (double)p.batch * nwin * (4.0 * S * (double)S * D_ + 2.0 * S * (double)(wsh + wsw) * D_),
                   (double)p.batch * gh * gw * 4.0 * D_ * (is_bf16 ? 2 : 4), stream);
    if (is_bf16) {
        SVB_CHECK_CUDA(cudaFuncSetAttribute(attention_simt_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attention_simt_kernel<bf16><<<grid, NT, smem, stream>>>(p);
    } else {
        SVB_CHECK_CUDA(cudaFuncSetAttribute(attention_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attention_simt_kernel<float><<<grid, NT, smem, stream>>>(p);
    }
    SVB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace svb
