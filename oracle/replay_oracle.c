/*
 * replay_oracle.c -- CPU restatement of border's replay-buffer path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (border_b200/, include/)
 * links, imports or executes this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, as the
 * checker.
 *
 * What it follows (all under /root/reference/):
 *   SimpleReplayBuffer   border-core/src/generic_replay_buffer/base.rs:86-426
 *   SumTree              border-core/src/generic_replay_buffer/base/sum_tree.rs:21-157
 *   IwScheduler          border-core/src/generic_replay_buffer/base/iw_scheduler.rs:6-46
 *   TensorBatch          border-tch-agent/src/tensor_batch.rs:85-120 (row store / index_select)
 *
 * Third-party algorithms that are NOT under /root/reference and are restated
 * from their published definitions:
 *   rand 0.8.5 StdRng = rand_chacha 0.3 ChaCha12Rng, seed_from_u64 = rand_core 0.6 PCG32 fill
 *   fastrand 1.x      = wyrand
 *   segment-tree 2.0.0 SegmentPoint<f32, Min/MaxIgnoreNaN>
 *   f32::powf         = host libm powf (called directly here)
 *
 * Parity pinning: SumTree::get is pinned by the reference's own known-answer
 * test (sum_tree.rs:180-216, tests/test_oracle_replay.py).  The ChaCha block
 * function is pinned by the published ChaCha20 (RFC 7539 2.3.2) and ChaCha12/8
 * zero-key vectors.  The seeded StdRng index stream, the wyrand stream and PER
 * weights have no reference-held vectors: "parity unpinned" for those streams.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ ChaCha */

#define ROTL32(v, n) (((v) << (n)) | ((v) >> (32 - (n))))
#define QR(a, b, c, d)                                                                      \
    a += b; d ^= a; d = ROTL32(d, 16);                                                      \
    c += d; b ^= c; b = ROTL32(b, 12);                                                      \
    a += b; d ^= a; d = ROTL32(d, 8);                                                       \
    c += d; b ^= c; b = ROTL32(b, 7);

/* One ChaCha block: `in` is the 16-word state, `rounds` is 8, 12 or 20. */
void bo_chacha_block(const uint32_t in[16], int rounds, uint32_t out[16]) {
    uint32_t x[16];
    memcpy(x, in, sizeof(x));
    for (int r = 0; r < rounds; r += 2) {
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13])
        QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12])
        QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
}

/* rand_chacha 0.3 ChaCha12Rng: key = seed (8 LE words), 64-bit block counter in
 * words 12-13, 64-bit stream id (0) in words 14-15; next_u32 consumes the
 * keystream words sequentially. */
typedef struct {
    uint32_t key[8];
    uint64_t word_pos; /* index of the next keystream word */
    uint32_t buf[16];
    uint64_t buf_block; /* block number held in buf, or UINT64_MAX */
} bo_stdrng;

/* rand_core 0.6 SeedableRng::seed_from_u64: PCG32 output function fills the seed. */
void bo_stdrng_seed_from_u64(bo_stdrng* r, uint64_t state) {
    const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
    for (int i = 0; i < 8; ++i) {
        state = state * MUL + INC;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        r->key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    }
    r->word_pos = 0;
    r->buf_block = UINT64_MAX;
}

static void stdrng_state(const uint32_t key[8], uint64_t block, uint32_t st[16]) {
    st[0] = 0x61707865; st[1] = 0x3320646e; st[2] = 0x79622d32; st[3] = 0x6b206574;
    for (int i = 0; i < 8; ++i) st[4 + i] = key[i];
    st[12] = (uint32_t)block; st[13] = (uint32_t)(block >> 32);
    st[14] = 0; st[15] = 0;
}

uint32_t bo_stdrng_next_u32(bo_stdrng* r) {
    uint64_t block = r->word_pos >> 4;
    if (block != r->buf_block) {
        uint32_t st[16];
        stdrng_state(r->key, block, st);
        bo_chacha_block(st, 12, r->buf);
        r->buf_block = block;
    }
    return r->buf[r->word_pos++ & 15];
}

/* ------------------------------------------------------------------ wyrand */

typedef struct { uint64_t s; } bo_fastrand;

void bo_fastrand_seed(bo_fastrand* r, uint64_t seed) { r->s = seed; }

uint64_t bo_fastrand_u64(bo_fastrand* r) {
    r->s += 0xA0761D6478BD642FULL;
    __uint128_t t = (__uint128_t)r->s * (__uint128_t)(r->s ^ 0xE7037ED1A0B428DBULL);
    return (uint64_t)t ^ (uint64_t)(t >> 64);
}
uint32_t bo_fastrand_u32(bo_fastrand* r) { return (uint32_t)bo_fastrand_u64(r); }

float bo_fastrand_f32(bo_fastrand* r) {
    uint32_t bits = 0x3F800000u | (bo_fastrand_u32(r) >> 9);
    float f; memcpy(&f, &bits, 4);
    return f - 1.0f;
}
double bo_fastrand_f64(bo_fastrand* r) {
    uint64_t bits = 0x3FF0000000000000ULL | (bo_fastrand_u64(r) >> 12);
    double f; memcpy(&f, &bits, 8);
    return f - 1.0;
}
/* fastrand u32(..n): Lemire multiply-high with rejection. */
uint32_t bo_fastrand_u32_below(bo_fastrand* r, uint32_t n) {
    uint32_t x = bo_fastrand_u32(r);
    uint64_t m = (uint64_t)x * n;
    uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
    if (lo < n) {
        uint32_t t = (0u - n) % n;
        while (lo < t) {
            x = bo_fastrand_u32(r);
            m = (uint64_t)x * n;
            hi = (uint32_t)(m >> 32); lo = (uint32_t)m;
        }
    }
    return hi;
}

/* ------------------------------------------------------ segment-tree 2.0.0 */
/* SegmentPoint: buf[2n], leaves at n.., buf[i] = op(buf[2i], buf[2i+1]). */

static float min_ignore_nan(float a, float b) { return isnan(a) ? b : (isnan(b) ? a : (a < b ? a : b)); }
static float max_ignore_nan(float a, float b) { return isnan(a) ? b : (isnan(b) ? a : (a > b ? a : b)); }

typedef struct { float* buf; size_t n; int is_max; } bo_segpoint;

static float seg_op(const bo_segpoint* s, float a, float b) {
    return s->is_max ? max_ignore_nan(a, b) : min_ignore_nan(a, b);
}
static void seg_build(bo_segpoint* s, size_t n, float fill, int is_max) {
    s->n = n; s->is_max = is_max;
    s->buf = (float*)malloc(sizeof(float) * 2 * n);
    for (size_t i = 0; i < 2 * n; ++i) s->buf[i] = fill;
    for (size_t i = n - 1; i >= 1; --i) s->buf[i] = seg_op(s, s->buf[2 * i], s->buf[2 * i + 1]);
}
static void seg_modify(bo_segpoint* s, size_t p, float v) {
    p += s->n; s->buf[p] = v;
    while (p > 1) { p >>= 1; s->buf[p] = seg_op(s, s->buf[2 * p], s->buf[2 * p + 1]); }
}
/* query over the half-open range [l, r) */
static float seg_query(const bo_segpoint* s, size_t l, size_t r) {
    float resl = s->is_max ? -INFINITY : INFINITY, resr = resl; /* identity of the op */
    int hasl = 0, hasr = 0;
    l += s->n; r += s->n;
    while (l < r) {
        if (l & 1) { resl = hasl ? seg_op(s, resl, s->buf[l]) : s->buf[l]; hasl = 1; l++; }
        if (r & 1) { r--; resr = hasr ? seg_op(s, s->buf[r], resr) : s->buf[r]; hasr = 1; }
        l >>= 1; r >>= 1;
    }
    if (hasl && hasr) return seg_op(s, resl, resr);
    return hasl ? resl : resr;
}

/* ------------------------------------------------------------------ SumTree */

typedef struct {
    float eps, alpha;
    size_t capacity, n_samples;
    float* tree; /* 2*capacity-1 */
    bo_segpoint min_tree, max_tree;
    int normalize; /* 0 = All, 1 = Batch */
} bo_sumtree;

bo_sumtree* bo_sumtree_new(size_t capacity, float alpha, int normalize) { /* sum_tree.rs:33-44 */
    bo_sumtree* t = (bo_sumtree*)calloc(1, sizeof(*t));
    t->eps = 1e-8f; t->alpha = alpha; t->capacity = capacity; t->n_samples = 0;
    t->tree = (float*)calloc(2 * capacity - 1, sizeof(float));
    seg_build(&t->min_tree, capacity, 3.40282347e+38f, 0);
    seg_build(&t->max_tree, capacity, 1e-8f, 1);
    t->normalize = normalize;
    return t;
}
void bo_sumtree_free(bo_sumtree* t) {
    if (!t) return;
    free(t->tree); free(t->min_tree.buf); free(t->max_tree.buf); free(t);
}
float bo_sumtree_total(const bo_sumtree* t) { return t->tree[0]; }
float bo_sumtree_max(const bo_sumtree* t) { /* sum_tree.rs:73-77 */
    return powf(seg_query(&t->max_tree, 0, t->capacity), 1.0f / t->alpha);
}
void bo_sumtree_update(bo_sumtree* t, size_t ix, float p) { /* sum_tree.rs:93-107 */
    p = powf(p + t->eps, t->alpha);
    seg_modify(&t->min_tree, ix, p);
    seg_modify(&t->max_tree, ix, p);
    size_t i = ix + t->capacity - 1;
    float change = p - t->tree[i];
    t->tree[i] = p;
    while (i != 0) { /* propagate, sum_tree.rs:46-52 */
        i = (i - 1) / 2;
        t->tree[i] += change;
    }
}
void bo_sumtree_add(bo_sumtree* t, size_t ix, float p) { /* sum_tree.rs:82-90 */
    bo_sumtree_update(t, ix, p);
    if (t->n_samples < t->capacity) t->n_samples++;
}
size_t bo_sumtree_get(const bo_sumtree* t, float s) { /* sum_tree.rs:54-67,110-114 */
    size_t ix = 0, len = 2 * t->capacity - 1;
    for (;;) {
        size_t left = 2 * ix + 1, right = left + 1;
        if (left >= len) break;
        if (s <= t->tree[left] || t->tree[right] == 0.0f) ix = left;
        else { s -= t->tree[left]; ix = right; }
    }
    return ix + 1 - t->capacity;
}
/* sum_tree.rs:120-157; `u` are the fastrand::f32() draws (injected). */
void bo_sumtree_sample(const bo_sumtree* t, size_t batch, float beta, const float* u,
                       int64_t* ixs, float* ws) {
    float p_sum = bo_sumtree_total(t);
    for (size_t k = 0; k < batch; ++k) ixs[k] = (int64_t)bo_sumtree_get(t, p_sum * u[k]);
    float n = (float)t->n_samples / p_sum;
    for (size_t k = 0; k < batch; ++k) ws[k] = powf(n * t->tree[ixs[k] + t->capacity - 1], -beta);
    float w_max_inv;
    if (t->normalize == 0) {
        w_max_inv = powf(n * seg_query(&t->min_tree, 0, t->n_samples), beta);
    } else {
        float m = NAN; /* fold(0.0/0.0, |m, v| v.max(m)) : f32::max ignores NaN */
        for (size_t k = 0; k < batch; ++k) m = fmaxf(ws[k], m);
        w_max_inv = 1.0f / m;
    }
    for (size_t k = 0; k < batch; ++k) ws[k] = ws[k] * w_max_inv;
}
const float* bo_sumtree_tree(const bo_sumtree* t) { return t->tree; }
size_t bo_sumtree_n_samples(const bo_sumtree* t) { return t->n_samples; }
float bo_sumtree_min_query(const bo_sumtree* t, size_t l, size_t r) { return seg_query(&t->min_tree, l, r); }

/* -------------------------------------------------------------- IwScheduler */

typedef struct { float beta_0, beta_final; size_t n_opts_final, n_opts; } bo_iw;

float bo_iw_beta(const bo_iw* s) { /* iw_scheduler.rs:33-41 */
    if (s->n_opts >= s->n_opts_final) return s->beta_final;
    float d = s->beta_final - s->beta_0;
    return s->beta_0 + d * ((float)s->n_opts / (float)s->n_opts_final);
}

/* ------------------------------------------------------- SimpleReplayBuffer */

typedef struct {
    size_t capacity, i, size;
    size_t obs_bytes, act_bytes; /* bytes per row (TensorBatch rows) */
    uint8_t *obs, *next_obs, *act;
    float* reward; int8_t *term, *trunc;
    bo_stdrng rng;
    bo_fastrand fr;
    int per; bo_sumtree* st; bo_iw iw;
} bo_replay;

bo_replay* bo_replay_build(size_t capacity, uint64_t seed, size_t obs_bytes, size_t act_bytes,
                           int per, float alpha, float beta_0, float beta_final,
                           size_t n_opts_final, int normalize, uint64_t fastrand_seed) {
    /* base.rs:336-356 */
    bo_replay* r = (bo_replay*)calloc(1, sizeof(*r));
    r->capacity = capacity; r->obs_bytes = obs_bytes; r->act_bytes = act_bytes;
    r->obs = (uint8_t*)calloc(capacity, obs_bytes);
    r->next_obs = (uint8_t*)calloc(capacity, obs_bytes);
    r->act = (uint8_t*)calloc(capacity, act_bytes);
    r->reward = (float*)calloc(capacity, sizeof(float));
    r->term = (int8_t*)calloc(capacity, 1);
    r->trunc = (int8_t*)calloc(capacity, 1);
    bo_stdrng_seed_from_u64(&r->rng, seed);
    bo_fastrand_seed(&r->fr, fastrand_seed);
    r->per = per;
    if (per) {
        r->st = bo_sumtree_new(capacity, alpha, normalize);
        r->iw.beta_0 = beta_0; r->iw.beta_final = beta_final;
        r->iw.n_opts_final = n_opts_final; r->iw.n_opts = 0;
    }
    return r;
}
void bo_replay_free(bo_replay* r) {
    if (!r) return;
    free(r->obs); free(r->next_obs); free(r->act); free(r->reward); free(r->term); free(r->trunc);
    bo_sumtree_free(r->st); free(r);
}
size_t bo_replay_len(const bo_replay* r) { return r->size; }
size_t bo_replay_head(const bo_replay* r) { return r->i; }

/* base.rs:295-316 (push), :227-235 (set_priority), tensor_batch.rs:104-108 */
void bo_replay_push(bo_replay* r, const void* obs, const void* act, const void* next_obs,
                    const float* reward, const int8_t* term, const int8_t* trunc, size_t len) {
    for (size_t j = 0; j < len; ++j) {
        size_t k = (r->i + j) % r->capacity;
        memcpy(r->obs + k * r->obs_bytes, (const uint8_t*)obs + j * r->obs_bytes, r->obs_bytes);
        memcpy(r->act + k * r->act_bytes, (const uint8_t*)act + j * r->act_bytes, r->act_bytes);
        memcpy(r->next_obs + k * r->obs_bytes, (const uint8_t*)next_obs + j * r->obs_bytes, r->obs_bytes);
        r->reward[k] = reward[j]; r->term[k] = term[j]; r->trunc[k] = trunc[j];
    }
    if (r->per) {
        float max_p = bo_sumtree_max(r->st);
        for (size_t j = 0; j < len; ++j) bo_sumtree_add(r->st, (r->i + j) % r->capacity, max_p);
    }
    r->i = (r->i + len) % r->capacity;
    r->size += len;
    if (r->size >= r->capacity) r->size = r->capacity;
}

/* base.rs:376-402.  Index generation only (the part that must be bit-exact).
 * `u_inject` (may be NULL) replaces the fastrand::f32() draws of the PER path. */
void bo_replay_sample_indices(bo_replay* r, size_t batch, const float* u_inject,
                              uint64_t* ixs, float* weight /* may be NULL unless PER */) {
    if (r->per) {
        float* u = (float*)malloc(sizeof(float) * batch);
        int64_t* ii = (int64_t*)malloc(sizeof(int64_t) * batch);
        for (size_t k = 0; k < batch; ++k) u[k] = u_inject ? u_inject[k] : bo_fastrand_f32(&r->fr);
        bo_sumtree_sample(r->st, batch, bo_iw_beta(&r->iw), u, ii, weight);
        for (size_t k = 0; k < batch; ++k) ixs[k] = (uint64_t)ii[k];
        free(u); free(ii);
    } else {
        for (size_t k = 0; k < batch; ++k) ixs[k] = (uint64_t)bo_stdrng_next_u32(&r->rng) % r->size;
    }
}
/* TensorBatch::sample = index_select(0, ixs) plus the Vec columns (base.rs:391-400). */
void bo_replay_gather(const bo_replay* r, const uint64_t* ixs, size_t batch, void* obs, void* act,
                      void* next_obs, float* reward, int8_t* term, int8_t* trunc) {
    for (size_t k = 0; k < batch; ++k) {
        size_t ix = ixs[k];
        if (obs) memcpy((uint8_t*)obs + k * r->obs_bytes, r->obs + ix * r->obs_bytes, r->obs_bytes);
        if (act) memcpy((uint8_t*)act + k * r->act_bytes, r->act + ix * r->act_bytes, r->act_bytes);
        if (next_obs) memcpy((uint8_t*)next_obs + k * r->obs_bytes, r->next_obs + ix * r->obs_bytes, r->obs_bytes);
        if (reward) reward[k] = r->reward[ix];
        if (term) term[k] = r->term[ix];
        if (trunc) trunc[k] = r->trunc[ix];
    }
}
/* base.rs:413-426 */
void bo_replay_update_priority(bo_replay* r, const uint64_t* ixs, const float* td, size_t n) {
    if (!r->per) return;
    for (size_t k = 0; k < n; ++k) bo_sumtree_update(r->st, ixs[k], td[k]);
    r->iw.n_opts++;
}
bo_sumtree* bo_replay_sumtree(bo_replay* r) { return r->st; }
float bo_replay_beta(const bo_replay* r) { return r->per ? bo_iw_beta(&r->iw) : 0.0f; }

/* Host powf, exported so tests can compare the device restatement against it. */
float bo_powf(float x, float y) { return powf(x, y); }
