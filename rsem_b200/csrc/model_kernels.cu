// K1: per-hit conditional probabilities (getConPrb / getNoiseConPrb) and
// K3: model sufficient statistics (update / updateNoise) for the four RSEM read models.
//
// Reference semantics (paths relative to /root/reference):
//   getConPrb        SingleModel.h:95-146, SingleQModel.h:101-151, PairedEndModel.h:90-134,
//                    PairedEndQModel.h:94-138
//   getNoiseConPrb   SingleModel.h:148-157, SingleQModel.h:153-162, PairedEndModel.h:136-151,
//                    PairedEndQModel.h:140-155
//   update/updateNoise  SingleModel.h:163-215, SingleQModel.h:168-221, PairedEndModel.h:156-183,
//                    PairedEndQModel.h:161-188  (+ Profile.h:91-96, QProfile.h:88-93,
//                    NoiseProfile.h:71-76, NoiseQProfile.h:74-79, LenDist.h:46-49, RSPD.h:43-59)
//
// B200 mapping: reads (base codes + qualities) and transcripts are resident in HBM, parsed once
// (the reference re-parses the FASTA/FASTQ text every round).  A read is handled by a group of G
// lanes (the same grouping as the E-step); each lane walks the bases of its own hits.  The
// (Q)Profile table is staged in shared memory.  K3 accumulates with red.global.add.f64 into
// kReplicas copies of the statistics block (the ~20 hot (quality, base, base) cells would
// otherwise serialise in one L2 slice) which a second kernel folds into copy 0.
#include <algorithm>

#include "common.cuh"

namespace rsem_b200 {
namespace {

constexpr int kReplicas = 64;
constexpr int kBlock = 256;

struct ModelArgs {
    DevModel m;
    // reads
    int n_mates;
    const unsigned long long* roff[2];
    const unsigned char* rbase[2];
    const unsigned char* rqual[2];
    const unsigned char* lowq;
    // refs
    const unsigned long long* seq_off;
    const unsigned char* seq;
    const int* full_len;
    const int* tot_len;
    const unsigned long long* mask_off;
    const unsigned int* mask_words;
    // hits
    unsigned long long N;
    const unsigned long long* row_ptr;
    const int* sid;
    const int* pos;
    const int* insertL;
    double* conprb;
    double* ncpv;
    const double* post;
    const double* post0;
    // K3 targets
    double* stats;          // kReplicas blocks of `stats_block` doubles
    size_t stats_block;
    size_t o_noise, o_gld, o_rspd;  // offsets inside a block (profile at 0)
    int gld_lb, gld_span;
    int prof_rows_smem;     // rows of the profile table staged in shared memory (0 = use global)
    int* err_flag;
};

__device__ __forceinline__ double ld_adj(const DevLenDist& d, int len, int refL) {
    if (len <= d.lb || len > d.ub || refL <= d.lb) return 0.0;
    const int top = min(d.ub, refL) - d.lb;
    return d.pdf[len - d.lb] / d.cdf[top];
}
__device__ __forceinline__ double ld_prob(const DevLenDist& d, int len) {
    if (len <= d.lb || len > d.ub) return 0.0;  // the reference asserts this range (LenDist.h:58)
    return d.pdf[len - d.lb];
}
__device__ __forceinline__ double rspd_cdf_at(const DevModel& m, int fpos, int fullLen) {
    const int i = (int)(((long long)fpos) * m.rspd_B / fullLen);
    const double val = __dmul_rn(__ddiv_rn(fpos * 1.0, (double)fullLen), (double)m.rspd_B);
    return __dadd_rn(m.rspd_cdf[i], __dmul_rn(val - i, m.rspd_pdf[i + 1]));
}
__device__ __forceinline__ double rspd_adj(const DevModel& m, int fpos, int effL, int fullLen) {
    if (!m.est_rspd) return 1.0 / effL;
    const double denom = rspd_cdf_at(m, effL, fullLen);
    return denom >= kEpsilon ? (rspd_cdf_at(m, fpos + 1, fullLen) - rspd_cdf_at(m, fpos, fullLen)) / denom : 0.0;
}
__device__ __forceinline__ bool ref_mask(const ModelArgs& a, int sid, int p) {
    return (a.mask_words[a.mask_off[sid] + (p >> 5)] >> (p & 31)) & 1u;
}

// product over the bases of one mate of p[row][ref][read]; row = quality (Q models) or position
template <bool HASQ>
__device__ __forceinline__ double seq_prob(const ModelArgs& a, const double* prof, int mate, unsigned long long i,
                                           int sid, int pos, int dir) {
    const unsigned long long o = a.roff[mate][i];
    const int len = (int)(a.roff[mate][i + 1] - o);
    const int totLen = a.tot_len[sid];
    if (pos < 0 || pos + len > totLen) {  // reference: general_assert -> exit(-1) (e.g. SingleQModel.h:116-121)
        *a.err_flag = 2;
        return 0.0;
    }
    const unsigned char* rb = a.rbase[mate] + o;
    const unsigned char* rq = HASQ ? a.rqual[mate] + o : nullptr;
    const unsigned char* s = a.seq + a.seq_off[sid] + (dir == 0 ? pos : totLen - 1 - pos);
    double prob = 1.0;
    if (dir == 0) {
        for (int k = 0; k < len; ++k) {
            const int row = HASQ ? rq[k] : k;
            prob *= prof[(row * 5 + s[k]) * 5 + rb[k]];
        }
    } else {
        for (int k = 0; k < len; ++k) {
            const int c = s[-k];
            const int row = HASQ ? rq[k] : k;
            prob *= prof[(row * 5 + (c < 4 ? 3 - c : 4)) * 5 + rb[k]];
        }
    }
    return prob;
}

template <bool HASQ>
__device__ __forceinline__ double noise_prob(const ModelArgs& a, const double* nprof, int mate, unsigned long long i) {
    const unsigned long long o = a.roff[mate][i];
    const int len = (int)(a.roff[mate][i + 1] - o);
    const unsigned char* rb = a.rbase[mate] + o;
    const unsigned char* rq = HASQ ? a.rqual[mate] + o : nullptr;
    double prob = 1.0;
    for (int k = 0; k < len; ++k) prob *= HASQ ? nprof[rq[k] * 5 + rb[k]] : nprof[rb[k]];
    return prob;
}

template <bool HASQ>
__device__ double conprb_single(const ModelArgs& a, const double* prof, unsigned long long i, int s, int pos) {
    const DevModel& m = a.m;
    const int sid = abs(s), dir = s < 0;
    const int fullLen = a.full_len[sid], totLen = a.tot_len[sid];
    const int readLen = (int)(a.roff[0][i + 1] - a.roff[0][i]);
    const int fpos = dir == 0 ? pos : totLen - pos - readLen;
    const int seedPos = dir == 0 ? pos : totLen - pos - m.seed_len;
    if (pos < 0 || fpos < 0 || seedPos < 0) { *a.err_flag = 2; return 0.0; }
    if (seedPos >= fullLen || ref_mask(a, sid, seedPos)) return 0.0;
    double value;
    if (m.has_mld) {
        const int minL = max(readLen, m.gld.lb + 1), maxL = min(totLen - pos, m.gld.ub);
        value = 0.0;
        for (int fragLen = minL; fragLen <= maxL; ++fragLen) {
            const int pfpos = dir == 0 ? pos : totLen - pos - fragLen;
            const int effL = min(fullLen, totLen - fragLen + 1);
            value += ld_adj(m.gld, fragLen, totLen) * rspd_adj(m, pfpos, effL, fullLen) * ld_adj(m.mld, readLen, fragLen);
        }
    } else {
        const int effL = min(fullLen, totLen - readLen + 1);
        value = ld_adj(m.gld, readLen, totLen) * rspd_adj(m, fpos, effL, fullLen);
    }
    double prob = m.ori[dir] * value * seq_prob<HASQ>(a, prof, 0, i, sid, pos, dir);
    if (prob < kEpsilon) prob = 0.0;
    const double w = m.mw[sid];
    return w < kEpsilon ? 0.0 : prob / w;
}

template <bool HASQ>
__device__ double conprb_paired(const ModelArgs& a, const double* prof, unsigned long long i, int s, int pos,
                                int insertLen) {
    const DevModel& m = a.m;
    const int sid = abs(s), dir = s < 0;
    const int fullLen = a.full_len[sid], totLen = a.tot_len[sid];
    const int fpos = dir == 0 ? pos : totLen - pos - insertLen;
    const int effL = min(fullLen, totLen - insertLen + 1);
    if (pos < 0 || fpos < 0 || insertLen > totLen) { *a.err_flag = 2; return 0.0; }
    if (fpos >= fullLen || ref_mask(a, sid, fpos)) return 0.0;
    double prob = m.ori[dir] * ld_adj(m.gld, insertLen, totLen) * rspd_adj(m, fpos, effL, fullLen);
    const int len1 = (int)(a.roff[0][i + 1] - a.roff[0][i]), len2 = (int)(a.roff[1][i + 1] - a.roff[1][i]);
    prob *= ld_adj(m.mld, len1, insertLen) * seq_prob<HASQ>(a, prof, 0, i, sid, pos, dir);
    prob *= ld_adj(m.mld, len2, insertLen) * seq_prob<HASQ>(a, prof, 1, i, sid, totLen - pos - insertLen, !dir);
    if (prob < kEpsilon) prob = 0.0;
    const double w = m.mw[sid];
    return w < kEpsilon ? 0.0 : prob / w;
}

template <bool HASQ>
__device__ double noise_conprb(const ModelArgs& a, const double* nprof, unsigned long long i) {
    const DevModel& m = a.m;
    double prob;
    if (m.model_type < 2) {
        const int readLen = (int)(a.roff[0][i + 1] - a.roff[0][i]);
        prob = m.has_mld ? ld_prob(m.mld, readLen) : ld_prob(m.gld, readLen);
        prob *= noise_prob<HASQ>(a, nprof, 0, i);
    } else {
        const int len1 = (int)(a.roff[0][i + 1] - a.roff[0][i]), len2 = (int)(a.roff[1][i + 1] - a.roff[1][i]);
        prob = ld_prob(m.mld, len1) * noise_prob<HASQ>(a, nprof, 0, i);
        prob *= ld_prob(m.mld, len2) * noise_prob<HASQ>(a, nprof, 1, i);
    }
    if (prob < kEpsilon) prob = 0.0;
    const double w = m.mw[0];
    return w < kEpsilon ? 0.0 : prob / w;
}

// stage the first `rows` rows (25 doubles each) of the profile + the noise table in shared memory
__device__ __forceinline__ void stage_tables(const ModelArgs& a, double* smem, const double*& prof, const double*& nprof,
                                             bool hasq) {
    prof = a.m.profile;
    nprof = a.m.noise_profile;
    if (a.prof_rows_smem > 0) {
        const int n = a.prof_rows_smem * 25;
        for (int k = threadIdx.x; k < n; k += blockDim.x) smem[k] = a.m.profile[k];
        const int nn = hasq ? 500 : 5;
        for (int k = threadIdx.x; k < nn; k += blockDim.x) smem[n + k] = a.m.noise_profile[k];
        __syncthreads();
        prof = smem;
        nprof = smem + n;
    }
}

// ---- K1 ---------------------------------------------------------------------------------------
template <int G, bool HASQ, bool PAIRED>
__global__ void __launch_bounds__(kBlock) conprb_kernel(const ModelArgs a) {
    extern __shared__ double smem_tab[];
    const double *prof, *nprof;
    stage_tables(a, smem_tab, prof, nprof, HASQ);
    const int lane = threadIdx.x & 31, g = lane % G;
    const unsigned long long groups_total = (unsigned long long)gridDim.x * (kBlock / G);
    for (unsigned long long i = (unsigned long long)blockIdx.x * (kBlock / G) + threadIdx.x / G; i < a.N; i += groups_total) {
        const unsigned long long fr = a.row_ptr[i], to = a.row_ptr[i + 1];
        const bool lq = a.lowq[i] != 0;
        if (g == 0) a.ncpv[i] = lq ? 0.0 : noise_conprb<HASQ>(a, nprof, i);
        for (unsigned long long j = fr + g; j < to; j += G) {
            double v = 0.0;
            if (!lq)
                v = PAIRED ? conprb_paired<HASQ>(a, prof, i, a.sid[j], a.pos[j], a.insertL[j])
                           : conprb_single<HASQ>(a, prof, i, a.sid[j], a.pos[j]);
            a.conprb[j] = v;
        }
    }
}

// ---- K3 ---------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add(double* p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

template <bool HASQ>
__device__ __forceinline__ void prof_update(const ModelArgs& a, double* tab, int mate, unsigned long long i, int sid,
                                            int pos, int dir, double frac) {
    const unsigned long long o = a.roff[mate][i];
    const int len = (int)(a.roff[mate][i + 1] - o);
    const int totLen = a.tot_len[sid];
    if (pos < 0 || pos + len > totLen) return;
    const unsigned char* rb = a.rbase[mate] + o;
    const unsigned char* rq = HASQ ? a.rqual[mate] + o : nullptr;
    const unsigned char* s = a.seq + a.seq_off[sid] + (dir == 0 ? pos : totLen - 1 - pos);
    for (int k = 0; k < len; ++k) {
        int c = dir == 0 ? s[k] : s[-k];
        if (dir) c = c < 4 ? 3 - c : 4;
        const int row = HASQ ? rq[k] : k;
        red_add(tab + (row * 5 + c) * 5 + rb[k], frac);
    }
}

template <bool HASQ>
__device__ __forceinline__ void noise_update(const ModelArgs& a, double* tab, int mate, unsigned long long i, double frac) {
    const unsigned long long o = a.roff[mate][i];
    const int len = (int)(a.roff[mate][i + 1] - o);
    const unsigned char* rb = a.rbase[mate] + o;
    const unsigned char* rq = HASQ ? a.rqual[mate] + o : nullptr;
    for (int k = 0; k < len; ++k) red_add(tab + (HASQ ? rq[k] * 5 + rb[k] : rb[k]), frac);
}

// RSPD::update, RSPD.h:43-59
__device__ __forceinline__ void rspd_update(const ModelArgs& a, double* pdf, int fpos, int fullLen, double frac) {
    if (fpos >= fullLen || fpos < 0) return;
    const int B = a.m.rspd_B;
    int i;
    double x = fpos * 1.0 / fullLen, b;
    for (i = (int)(((long long)fpos) * B / fullLen + 1); i < (((long long)fpos + 1) * B - 1) / fullLen + 1; ++i) {
        b = i * 1.0 / B;
        red_add(pdf + i, (b - x) * fullLen * frac);
        x = b;
    }
    b = (fpos + 1.0) / fullLen;
    red_add(pdf + i, (b - x) * fullLen * frac);
}

template <int G, bool HASQ, bool PAIRED>
__global__ void __launch_bounds__(kBlock) update_kernel(const ModelArgs a) {
    const int lane = threadIdx.x & 31, g = lane % G;
    const unsigned warp_global = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    double* blk = a.stats + (size_t)(warp_global % kReplicas) * a.stats_block;
    double* t_prof = blk;
    double* t_noise = blk + a.o_noise;
    double* t_gld = blk + a.o_gld;
    double* t_rspd = blk + a.o_rspd;
    const DevModel& m = a.m;
    const unsigned long long groups_total = (unsigned long long)gridDim.x * (kBlock / G);
    for (unsigned long long i = (unsigned long long)blockIdx.x * (kBlock / G) + threadIdx.x / G; i < a.N; i += groups_total) {
        if (a.lowq[i]) continue;
        const unsigned long long fr = a.row_ptr[i], to = a.row_ptr[i + 1];
        if (g == 0) {
            const double f0 = a.post0[i];
            if (f0 >= kEpsilon) {
                noise_update<HASQ>(a, t_noise, 0, i, f0);
                if (PAIRED) noise_update<HASQ>(a, t_noise, 1, i, f0);
            }
        }
        for (unsigned long long j = fr + g; j < to; j += G) {
            const double frac = a.post[j];
            if (frac < kEpsilon) continue;
            const int s = a.sid[j], t = abs(s), dir = s < 0, p = a.pos[j];
            const int fullLen = a.full_len[t], totLen = a.tot_len[t];
            if (!PAIRED) {
                if (m.est_rspd) {  // one strand only (SingleQModel.h:180-184; helper models have mld == NULL)
                    const int readLen = (int)(a.roff[0][i + 1] - a.roff[0][i]);
                    if (m.ori[0] >= 0.1 && dir == 0) rspd_update(a, t_rspd, p, fullLen, frac);
                    if (m.ori[0] < 0.1 && dir == 1) rspd_update(a, t_rspd, totLen - p - readLen, fullLen, frac);
                }
                prof_update<HASQ>(a, t_prof, 0, i, t, p, dir, frac);
            } else {
                const int il = a.insertL[j];
                if (il > a.gld_lb && il <= a.gld_lb + a.gld_span) red_add(t_gld + (il - a.gld_lb), frac);
                if (m.est_rspd) rspd_update(a, t_rspd, dir == 0 ? p : totLen - p - il, fullLen, frac);
                prof_update<HASQ>(a, t_prof, 0, i, t, p, dir, frac);
                prof_update<HASQ>(a, t_prof, 1, i, t, totLen - p - il, !dir, frac);
            }
        }
    }
}

__global__ void fold_replicas_kernel(double* stats, size_t block) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= block) return;
    double s = 0.0;
    for (int r = 0; r < kReplicas; ++r) s += stats[(size_t)r * block + k];
    stats[k] = s;
}

void fill_args(rsem_b200_ctx* c, ModelArgs& a) {
    a.m = c->model;
    a.n_mates = c->reads.n_mates;
    for (int k = 0; k < 2; ++k) {
        a.roff[k] = reinterpret_cast<const unsigned long long*>(c->reads.off[k]);
        a.rbase[k] = c->reads.base[k];
        a.rqual[k] = c->reads.qual[k];
    }
    a.lowq = c->reads.lowq;
    a.seq_off = reinterpret_cast<const unsigned long long*>(c->refs.seq_off);
    a.seq = c->refs.seq;
    a.full_len = c->refs.full_len;
    a.tot_len = c->refs.tot_len;
    a.mask_off = reinterpret_cast<const unsigned long long*>(c->refs.mask_off);
    a.mask_words = c->refs.mask_words;
    a.N = c->N;
    a.row_ptr = reinterpret_cast<const unsigned long long*>(c->row_ptr);
    a.sid = c->sid;
    a.pos = c->pos;
    a.insertL = c->insertL;
    a.conprb = c->conprb;
    a.ncpv = c->ncpv;
    a.post = c->post;
    a.post0 = c->post0;
    a.err_flag = c->err_flag;
    const bool hasq = c->model.model_type & 1;
    int rows = hasq ? 100 : std::min(c->model.pro_len, std::max(c->reads.max_len, 1));
    if ((size_t)rows * 200 + 4096 > 200 * 1024) rows = 0;
    a.prof_rows_smem = rows;
    a.stats = c->stats_buf;
    a.stats_block = c->stats.total_doubles;
    a.o_noise = c->stats.noise_profile ? (size_t)(c->stats.noise_profile - c->stats_buf) : 0;
    a.o_gld = c->stats.gld_pdf ? (size_t)(c->stats.gld_pdf - c->stats_buf) : 0;
    a.o_rspd = c->stats.rspd_pdf ? (size_t)(c->stats.rspd_pdf - c->stats_buf) : 0;
    a.gld_lb = c->stats.gld_lb;
    a.gld_span = c->stats.gld_span;
}

template <int G>
int launch_conprb_g(rsem_b200_ctx* c, const ModelArgs& a, unsigned grid, size_t smem) {
    const bool hasq = c->model.model_type & 1, paired = c->model.model_type >= 2;
#define RB_LAUNCH(Q, P)                                                                                      \
    do {                                                                                                     \
        auto k = conprb_kernel<G, Q, P>;                                                                     \
        if (smem > 48 * 1024) RB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k<<<grid, kBlock, smem, c->stream>>>(a);                                                             \
    } while (0)
    if (hasq && paired) RB_LAUNCH(true, true);
    else if (hasq) RB_LAUNCH(true, false);
    else if (paired) RB_LAUNCH(false, true);
    else RB_LAUNCH(false, false);
#undef RB_LAUNCH
    RB_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

template <int G>
int launch_update_g(rsem_b200_ctx* c, const ModelArgs& a, unsigned grid) {
    const bool hasq = c->model.model_type & 1, paired = c->model.model_type >= 2;
    if (hasq && paired) update_kernel<G, true, true><<<grid, kBlock, 0, c->stream>>>(a);
    else if (hasq) update_kernel<G, true, false><<<grid, kBlock, 0, c->stream>>>(a);
    else if (paired) update_kernel<G, false, true><<<grid, kBlock, 0, c->stream>>>(a);
    else update_kernel<G, false, false><<<grid, kBlock, 0, c->stream>>>(a);
    RB_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

unsigned grid_for(rsem_b200_ctx* c, int G) {
    const unsigned long long per_block = kBlock / G;
    unsigned long long want = (c->N + per_block - 1) / per_block;
    if (want < 1) want = 1;
    return (unsigned)std::min<unsigned long long>(want, (unsigned long long)c->sm_count * 8);
}

}  // namespace

int model_launch_conprb(rsem_b200_ctx* c) {
    if (c->N == 0) return 0;
    ModelArgs a;
    fill_args(c, a);
    const bool hasq = c->model.model_type & 1;
    const size_t smem = a.prof_rows_smem ? ((size_t)a.prof_rows_smem * 25 + (hasq ? 500 : 5)) * sizeof(double) : 0;
    const int G = c->group;
    const unsigned grid = grid_for(c, G);
    switch (G) {
        case 4: return launch_conprb_g<4>(c, a, grid, smem);
        case 8: return launch_conprb_g<8>(c, a, grid, smem);
        case 16: return launch_conprb_g<16>(c, a, grid, smem);
        default: return launch_conprb_g<32>(c, a, grid, smem);
    }
}

int model_launch_update(rsem_b200_ctx* c) {
    // (re)build the statistics block layout: [profile][noise][gld][rspd], kReplicas copies
    const bool hasq = c->model.model_type & 1, paired = c->model.model_type >= 2;
    const size_t n_prof = hasq ? 2500 : (size_t)c->model.pro_len * 25, n_noise = hasq ? 500 : 5;
    const size_t n_gld = paired ? (size_t)c->stats.gld_span + 1 : 0;
    const size_t n_rspd = c->model.est_rspd ? (size_t)c->model.rspd_B + 2 : 0;
    const size_t block = n_prof + n_noise + n_gld + n_rspd;
    if (c->stats_buf_doubles < block * kReplicas) {
        if (c->stats_buf) cudaFree(c->stats_buf);
        RB_CUDA(cudaMalloc(&c->stats_buf, block * kReplicas * sizeof(double)));
        c->stats_buf_doubles = block * kReplicas;
    }
    c->stats.total_doubles = block;
    c->stats.profile = c->stats_buf;
    c->stats.noise_profile = c->stats_buf + n_prof;
    c->stats.gld_pdf = c->stats_buf + n_prof + n_noise;
    c->stats.rspd_pdf = c->stats_buf + n_prof + n_noise + n_gld;
    RB_CUDA(cudaMemsetAsync(c->stats_buf, 0, block * kReplicas * sizeof(double), c->stream));
    if (c->N == 0) return 0;
    ModelArgs a;
    fill_args(c, a);
    const int G = c->group;
    const unsigned grid = grid_for(c, G);
    int rc;
    switch (G) {
        case 4: rc = launch_update_g<4>(c, a, grid); break;
        case 8: rc = launch_update_g<8>(c, a, grid); break;
        case 16: rc = launch_update_g<16>(c, a, grid); break;
        default: rc = launch_update_g<32>(c, a, grid); break;
    }
    if (rc) return rc;
    fold_replicas_kernel<<<(unsigned)((block + 255) / 256), 256, 0, c->stream>>>(c->stats_buf, block);
    RB_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

}  // namespace rsem_b200
