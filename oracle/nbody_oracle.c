/*
 * nbody_oracle.c -- CPU restatement of blitzcode/rust-exp rs-src/nbody.rs (the N-body hot path).
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the checker, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * Nothing under rust_exp_b200/ links, imports or calls it.
 *
 * PARITY STATUS: **parity unpinned by reference vectors**.  The reference ships no tests, golden
 * vectors or fixtures for this path and cannot be compiled in this image
 * (Rust, no cargo/rustc).  The restatement is therefore pinned only by (a) known-answer tests derived
 * by hand from the source semantics (tests/test_oracle_kat.py) and (b) an independent
 * numpy-float32 restatement (oracle/oracle_np.py) that must agree with it bit for bit.
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 * Arithmetic is plain IEEE binary32, one rounding per operation, no contraction:
 * build with  gcc -O2 -ffp-contract=off  on x86-64 (SSE2 scalar float, FLT_EVAL_METHOD == 0).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__FAST_MATH__)
#error "the oracle must not be built with -ffast-math"
#endif

/* rs-src/nbody.rs:13-17 */
static const float VP_WDH = 100.0f;
static const float VP_ORG_X = 0.0f;
static const float VP_ORG_Y = 0.0f;
static const float EPS = 0.0001f;

/* rs-src/nbody.rs:19-26 -- AoS record, 5 x f32, 20 bytes */
typedef struct {
    float px, py, vx, vy, m;
} Particle;

/* rs-src/nbody.rs:28-32 -- process-global particle vector (the mutex is the caller's business here:
 * the oracle is driven from one thread) */
static Particle *g_p = NULL;
static int32_t g_n = 0;
static int32_t g_cap = 0;

static void reserve(int32_t n)
{
    if (n > g_cap) {
        g_p = (Particle *)realloc(g_p, (size_t)n * sizeof(Particle));
        if (!g_p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
        g_cap = n;
    }
}

/* rs-src/nbody.rs:34-37 */
int32_t ora_num_particles(void) { return g_n; }

/* Additions: state injection / read-back, 20-byte AoS records. */
void ora_set_particles(const float *aos5, int32_t n)
{
    reserve(n);
    memcpy(g_p, aos5, (size_t)n * sizeof(Particle));
    g_n = n;
}
void ora_get_particles(float *aos5_out, int32_t n)
{
    if (n > g_n) n = g_n;
    memcpy(aos5_out, g_p, (size_t)n * sizeof(Particle));
}

/* ---------------------------------------------------------------------------------------------
 * Initial conditions.  rs-src/nbody.rs:39-104 draw from an UNSEEDED rand::thread_rng(), so only the
 * distribution can be matched; these use a seeded splitmix64 so that runs are reproducible.
 * ------------------------------------------------------------------------------------------- */
static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
void ora_seed(uint64_t s) { g_rng = s ? s : 0x9E3779B97F4A7C15ull; }
static uint64_t splitmix64(void)
{
    uint64_t z = (g_rng += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* uniform f32 in [0,1), 24 random bits */
static float u01(void) { return (float)(splitmix64() >> 40) * (1.0f / 16777216.0f); }
static float urange(float lo, float hi) { return lo + (hi - lo) * u01(); }

/* rs-src/nbody.rs:66-71 */
static void uniform_sample_disk(float *x, float *y)
{
    float r = sqrtf(*x);
    float theta = 2.0f * 3.14159265358979323846f * (*y);
    *x = r * cosf(theta);
    *y = r * sinf(theta);
}

/* rs-src/nbody.rs:39-64 */
void ora_random_disk(int32_t n)
{
    reserve(n);
    g_n = 0;
    for (int32_t i = 0; i < n; i++) {
        float x = u01(), y = u01();
        uniform_sample_disk(&x, &y);
        x *= 23.0f;
        y *= 23.0f;
        Particle p;
        p.px = x;
        p.py = y;
        p.vx = urange(-3.5f, 3.5f);
        p.vy = urange(-3.5f, 3.5f);
        p.m = urange(0.1f, 1.5f);
        g_p[g_n++] = p;
    }
}

/* rs-src/nbody.rs:73-104 */
void ora_stable_orbits(int32_t n, float rmin, float rmax)
{
    if (n < 1) { g_n = 0; return; }
    reserve(n);
    g_n = 0;
    const float sun_mass = 1000.0f, planet_mass = 1.0f, g = 1.0f;
    const float speed = sqrtf(g * sun_mass);
    Particle sun = {0.0f, 0.0f, 0.0f, 0.0f, sun_mass};
    g_p[g_n++] = sun;
    for (int32_t i = 0; i < n - 1; i++) {
        float r = (rmax - rmin) * u01() + rmin;
        float theta = 2.0f * 3.14159265358979323846f * u01();
        Particle p;
        p.px = r * cosf(theta);
        p.py = r * sinf(theta);
        p.vx = -speed * sinf(theta);
        p.vy = speed * cosf(theta);
        p.m = planet_mass;
        g_p[g_n++] = p;
    }
}

/* ---------------------------------------------------------------------------------------------
 * Pair law.  rs-src/nbody.rs:164-184.  NOTE: the direction vector is NOT normalised;
 * |F| ~ m1 m2 / r.  Operation order is the reference's: (m1*m2) / (d2 + EPS), then * dx, * dy.
 * ------------------------------------------------------------------------------------------- */
static inline void force(float px1, float py1, float m1, float px2, float py2, float m2,
                         float *fx, float *fy)
{
    float dx = px2 - px1;
    float dy = py2 - py1;
    float dist_sq = dx * dx + dy * dy;
    float f = m1 * m2 / (dist_sq + EPS);
    *fx = f * dx;
    *fy = f * dy;
}
void ora_force(float px1, float py1, float m1, float px2, float py2, float m2, float *out2)
{
    force(px1, py1, m1, px2, py2, m2, &out2[0], &out2[1]);
}

/* rs-src/nbody.rs:129-144 -- forces on rows [i0, i1), sequential f32 accumulation over ascending j.
 * Exposed so the CPU baseline can time a bounded sample of rows of a large system. */
void ora_brute_forces_rows(int32_t i0, int32_t i1, float *fxy_out /* 2*(i1-i0) */)
{
    for (int32_t i = i0; i < i1; i++) {
        const Particle *a = &g_p[i];
        float fx = 0.0f, fy = 0.0f;
        for (int32_t j = 0; j < g_n; j++) {
            if (i == j) continue;
            const Particle *b = &g_p[j];
            float fxa, fya;
            force(a->px, a->py, a->m, b->px, b->py, b->m, &fxa, &fya);
            fx += fxa;
            fy += fya;
        }
        fxy_out[2 * (i - i0) + 0] = fx;
        fxy_out[2 * (i - i0) + 1] = fy;
    }
}

/* rs-src/nbody.rs:153-160 -- semi-implicit Euler: v += (dt*F)/m ; p += dt*v (new v) */
static inline void euler(Particle *p, float fx, float fy, float dt)
{
    p->vx += dt * fx / p->m;
    p->vy += dt * fy / p->m;
    p->px += dt * p->vx;
    p->py += dt * p->vy;
}

/* rs-src/nbody.rs:106-162 */
void ora_step_brute_force(float dt)
{
    if (g_n == 0) return;
    float *f = (float *)malloc((size_t)g_n * 2 * sizeof(float));
    ora_brute_forces_rows(0, g_n, f);
    for (int32_t i = 0; i < g_n; i++) euler(&g_p[i], f[2 * i], f[2 * i + 1], dt);
    free(f);
}

/* NOT IN THE REFERENCE (rs-src/nbody.rs:132-144 is single-threaded): the same step with the i loop
 * split over nthreads.  Results are identical (rows are independent); reported only as a labelled
 * "all cores" CPU figure. */
typedef struct { int32_t i0, i1; float *f; } RowJob;
static void *row_worker(void *arg)
{
    RowJob *j = (RowJob *)arg;
    ora_brute_forces_rows(j->i0, j->i1, j->f + 2 * (size_t)j->i0);
    return NULL;
}
void ora_brute_forces_rows_mt(int32_t i0, int32_t i1, float *fxy_out, int32_t nthreads)
{
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    RowJob *jobs = (RowJob *)malloc(sizeof(RowJob) * nthreads);
    int32_t n = i1 - i0, per = (n + nthreads - 1) / nthreads;
    for (int32_t t = 0; t < nthreads; t++) {
        jobs[t].i0 = i0 + (t * per < n ? t * per : n);
        jobs[t].i1 = i0 + ((t + 1) * per < n ? (t + 1) * per : n);
        jobs[t].f = fxy_out - 2 * (size_t)i0;
        pthread_create(&th[t], NULL, row_worker, &jobs[t]);
    }
    for (int32_t t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

/* The same step as ora_step_brute_force with the force rows split over host threads (rows are independent
 * and every row still accumulates over ascending j, so the result is bit-identical); lets the checker
 * finish a 65,536-body step in well under a second.  NOT a reference code path. */
void ora_step_brute_force_mt(float dt, int32_t nthreads)
{
    if (g_n == 0) return;
    float *f = (float *)malloc((size_t)g_n * 2 * sizeof(float));
    ora_brute_forces_rows_mt(0, g_n, f, nthreads);
    for (int32_t i = 0; i < g_n; i++) euler(&g_p[i], f[2 * i], f[2 * i + 1], dt);
    free(f);
}

/* ---------------------------------------------------------------------------------------------
 * Barnes-Hut.  rs-src/nbody.rs:186-480.
 * ------------------------------------------------------------------------------------------- */

/* rs-src/nbody.rs:206-214.  children == NULL <=> Option::None; otherwise a block of 4 in the order
 * UL, UR, LL, LR (rs-src/nbody.rs:295-300). */
typedef struct Node {
    float x1, y1, x2, y2;
    float px, py, m;
    struct Node *children;
} Node;

/* Node blocks come from an arena instead of Box::new per split; allocation strategy has no effect
 * on results. */
typedef struct Arena {
    Node *blocks;
    size_t used, cap; /* in nodes */
    struct Arena *next;
} Arena;
static Arena *g_arena = NULL;
static size_t g_nodes_allocated = 0;

static Node *arena_alloc4(void)
{
    if (!g_arena || g_arena->used + 4 > g_arena->cap) {
        Arena *a = (Arena *)malloc(sizeof(Arena));
        a->cap = (size_t)1 << 18;
        a->used = 0;
        a->blocks = (Node *)malloc(a->cap * sizeof(Node));
        a->next = g_arena;
        g_arena = a;
    }
    Node *r = g_arena->blocks + g_arena->used;
    g_arena->used += 4;
    g_nodes_allocated += 4;
    return r;
}
static void arena_reset(void)
{
    while (g_arena) {
        Arena *n = g_arena->next;
        free(g_arena->blocks);
        free(g_arena);
        g_arena = n;
    }
    g_nodes_allocated = 0;
}

/* rs-src/nbody.rs:216-222 */
static void node_new(Node *n, float x1, float y1, float x2, float y2)
{
    n->x1 = x1; n->y1 = y1; n->x2 = x2; n->y2 = y2;
    n->px = 0.0f; n->py = 0.0f; n->m = 0.0f;
    n->children = NULL;
}

/* rs-src/nbody.rs:303-320 */
static void add_mass(Node *n, float px, float py, float m)
{
    if (!(m > 0.0f)) { fprintf(stderr, "oracle: assertion failed: m > 0.0\n"); abort(); }
    if (n->m == 0.0f) {
        n->px = px; n->py = py; n->m = m;
    } else {
        float inv_msum = 1.0f / (n->m + m);
        n->px = (n->px * n->m + px * m) * inv_msum;
        n->py = (n->py * n->m + py * m) * inv_msum;
        n->m += m;
    }
}

/* rs-src/nbody.rs:322-331 -- returns the child index: UL=0, UR=1, LL=2, LR=3 */
static int quadrant_from_point(const Node *n, float x, float y)
{
    float cx = (n->x1 + n->x2) * 0.5f;
    float cy = (n->y1 + n->y2) * 0.5f;
    if (y < cy) return (x < cx) ? 2 : 3;
    return (x < cx) ? 0 : 1;
}

/* rs-src/nbody.rs:286-301 */
static void create_children(Node *n)
{
    float cx = (n->x1 + n->x2) * 0.5f;
    float cy = (n->y1 + n->y2) * 0.5f;
    Node *c = arena_alloc4();
    node_new(&c[0], n->x1, cy, cx, n->y2);    /* UL */
    node_new(&c[1], cx, cy, n->x2, n->y2);    /* UR */
    node_new(&c[2], n->x1, n->y1, cx, cy);    /* LL */
    node_new(&c[3], cx, n->y1, n->x2, cy);    /* LR */
    n->children = c;
}

static int g_max_depth_seen = 0;

/* rs-src/nbody.rs:226-284 */
static void insert(Node *n, float px, float py, float m, unsigned depth)
{
    if (depth > 50) {
        fprintf(stderr, "oracle: Node::insert() - recursion px:%g py:%g m:%g\n", px, py, m);
        abort();
    }
    if ((int)depth > g_max_depth_seen) g_max_depth_seen = (int)depth;
    if (n->children) {
        add_mass(n, px, py, m);
        int q = quadrant_from_point(n, px, py);
        insert(&n->children[q], px, py, m, depth + 1);
    } else {
        int too_close = fabsf(n->px - px) < EPS && fabsf(n->py - py) < EPS;
        if (n->m == 0.0f || too_close) {
            add_mass(n, px, py, m);
        } else {
            float px_o = n->px, py_o = n->py, m_o = n->m;
            n->px = 0.0f; n->py = 0.0f; n->m = 0.0f;
            create_children(n);
            insert(n, px_o, py_o, m_o, depth + 1);
            insert(n, px, py, m, depth + 1);
        }
    }
}

static uint64_t g_count_interactions = 0; /* only touched by the single-threaded counting variant */

/* rs-src/nbody.rs:333-377.  The nested (tree-shaped) summation order is part of the semantics:
 * an opened node returns ((((0+c0)+c1)+c2)+c3). */
static void compute_force(const Node *n, float px, float py, float m, float theta,
                          float *ofx, float *ofy)
{
    float fx = 0.0f, fy = 0.0f;
    if (n->children) {
        float s = n->x2 - n->x1;
        float dx = n->px - px;
        float dy = n->py - py;
        float d = sqrtf(dx * dx + dy * dy);
        if (s / d < theta) {
            force(px, py, m, n->px, n->py, n->m, &fx, &fy);
        } else {
            for (int i = 0; i < 4; i++) {
                float fxa, fya;
                compute_force(&n->children[i], px, py, m, theta, &fxa, &fya);
                fx += fxa;
                fy += fya;
            }
        }
    } else {
        if (n->px == px && n->py == py) { *ofx = 0.0f; *ofy = 0.0f; return; }
        if (n->m == 0.0f) { *ofx = 0.0f; *ofy = 0.0f; return; }
        force(px, py, m, n->px, n->py, n->m, &fx, &fy);
    }
    *ofx = fx;
    *ofy = fy;
}

/* Same walk, counting force() evaluations and visited nodes (instrumentation for the byte/flop
 * accounting in DESIGN.md; not part of the reference). */
static void count_walk(const Node *n, float px, float py, float theta, uint64_t *inter, uint64_t *visited)
{
    (*visited)++;
    if (n->children) {
        float s = n->x2 - n->x1;
        float dx = n->px - px, dy = n->py - py;
        float d = sqrtf(dx * dx + dy * dy);
        if (s / d < theta) (*inter)++;
        else for (int i = 0; i < 4; i++) count_walk(&n->children[i], px, py, theta, inter, visited);
    } else {
        if (n->px == px && n->py == py) return;
        if (n->m == 0.0f) return;
        (*inter)++;
    }
}

static Node g_root;
static int g_tree_valid = 0;
static int g_square_aabb = 0;
void ora_set_square_aabb(int32_t on) { g_square_aabb = on != 0; }

/* rs-src/nbody.rs:388-417 -- tight, non-square AABB, then insertion in particle index order */
void ora_bh_build(void)
{
    arena_reset();
    g_max_depth_seen = 0;
    float x1 = 3.40282347e+38f, y1 = 3.40282347e+38f;
    float x2 = -3.40282347e+38f, y2 = -3.40282347e+38f; /* f32::MIN is the most negative finite */
    for (int32_t i = 0; i < g_n; i++) {
        const Particle *p = &g_p[i];
        x1 = p->px < x1 ? p->px : x1;
        y1 = p->py < y1 ? p->py : y1;
        x2 = p->px > x2 ? p->px : x2;
        y2 = p->py > y2 ? p->py : y2;
    }
    if (g_square_aabb) {
        /* NOT the reference's behaviour: its author's commented-out variant, rs-src/nbody.rs:400-407, kept as an
         * opt-in so that the product's nbx_set_square_aabb(1) has something to be checked against */
        if (x2 - x1 > y2 - y1) y2 = y1 + (x2 - x1);
        else x2 = x1 + (y2 - y1);
    }
    node_new(&g_root, x1, y1, x2, y2);
    for (int32_t i = 0; i < g_n; i++) insert(&g_root, g_p[i].px, g_p[i].py, g_p[i].m, 0);
    g_tree_valid = 1;
}

int32_t ora_bh_node_count(void) { return g_tree_valid ? (int32_t)(1 + g_nodes_allocated) : 0; }
int32_t ora_bh_max_depth(void) { return g_max_depth_seen; }

/* Flatten the current tree in DFS pre-order (children in index order).  Each record is 9 floats:
 * x1,y1,x2,y2,px,py,m,has_children(0/1),depth.  Returns the number of records written. */
static int32_t flatten(const Node *n, int depth, float *out, int32_t cap, int32_t at)
{
    if (at < cap) {
        float *r = out + 9 * (size_t)at;
        r[0] = n->x1; r[1] = n->y1; r[2] = n->x2; r[3] = n->y2;
        r[4] = n->px; r[5] = n->py; r[6] = n->m;
        r[7] = n->children ? 1.0f : 0.0f;
        r[8] = (float)depth;
    }
    at++;
    if (n->children)
        for (int i = 0; i < 4; i++) at = flatten(&n->children[i], depth + 1, out, cap, at);
    return at;
}
int32_t ora_bh_flatten(float *out9, int32_t cap)
{
    if (!g_tree_valid) return 0;
    return flatten(&g_root, 0, out9, cap, 0);
}

/* Forces from the current tree for rows [i0,i1) (rs-src/nbody.rs:447), no state update. */
void ora_bh_forces_rows(float theta, int32_t i0, int32_t i1, float *fxy_out)
{
    for (int32_t i = i0; i < i1; i++) {
        compute_force(&g_root, g_p[i].px, g_p[i].py, g_p[i].m, theta,
                      &fxy_out[2 * (i - i0)], &fxy_out[2 * (i - i0) + 1]);
    }
}

/* rows are independent: the same forces with the row loop split over host threads (checker speed only) */
typedef struct { int32_t lo, hi; float theta; float *out; int32_t i0; } BhRowJob;
static void *bh_row_worker(void *arg)
{
    BhRowJob *j = (BhRowJob *)arg;
    if (j->hi > j->lo) ora_bh_forces_rows(j->theta, j->lo, j->hi, j->out + 2 * (size_t)(j->lo - j->i0));
    return NULL;
}
void ora_bh_forces_rows_mt(float theta, int32_t i0, int32_t i1, float *fxy_out, int32_t nthreads)
{
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    BhRowJob *jobs = (BhRowJob *)malloc(sizeof(BhRowJob) * nthreads);
    int32_t n = i1 - i0, per = (n + nthreads - 1) / nthreads;
    for (int32_t t = 0; t < nthreads; t++) {
        jobs[t].lo = i0 + (t * per < n ? t * per : n);
        jobs[t].hi = i0 + ((t + 1) * per < n ? (t + 1) * per : n);
        jobs[t].theta = theta; jobs[t].out = fxy_out; jobs[t].i0 = i0;
        pthread_create(&th[t], NULL, bh_row_worker, &jobs[t]);
    }
    for (int32_t t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

typedef struct { int32_t lo, hi; float theta; uint64_t inter, visited; } CountJob;
static void *count_worker(void *arg)
{
    CountJob *j = (CountJob *)arg;
    uint64_t a = 0, b = 0;
    for (int32_t i = j->lo; i < j->hi; i++) count_walk(&g_root, g_p[i].px, g_p[i].py, j->theta, &a, &b);
    j->inter = a;
    j->visited = b;
    return NULL;
}
/* bodies are independent and the counts are integers: splitting the loop over threads changes nothing */
void ora_bh_count_mt(float theta, uint64_t *interactions, uint64_t *visited, int32_t nthreads)
{
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    CountJob *jobs = (CountJob *)malloc(sizeof(CountJob) * nthreads);
    int32_t per = (g_n + nthreads - 1) / nthreads;
    for (int32_t t = 0; t < nthreads; t++) {
        jobs[t].lo = t * per < g_n ? t * per : g_n;
        jobs[t].hi = (t + 1) * per < g_n ? (t + 1) * per : g_n;
        jobs[t].theta = theta;
        pthread_create(&th[t], NULL, count_worker, &jobs[t]);
    }
    uint64_t a = 0, b = 0;
    for (int32_t t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        a += jobs[t].inter;
        b += jobs[t].visited;
    }
    free(th);
    free(jobs);
    *interactions = a;
    *visited = b;
    g_count_interactions = a;
}
void ora_bh_count(float theta, uint64_t *interactions, uint64_t *visited)
{
    ora_bh_count_mt(theta, interactions, visited, 1);
}

typedef struct { int32_t lo, hi; float theta, dt; } BhJob;

/* rs-src/nbody.rs:443-472 -- per-body: tree force, Euler, velocity kill */
static void *bh_worker(void *arg)
{
    BhJob *j = (BhJob *)arg;
    for (int32_t i = j->lo; i < j->hi; i++) {
        Particle *p = &g_p[i];
        float fx, fy;
        compute_force(&g_root, p->px, p->py, p->m, j->theta, &fx, &fy);
        euler(p, fx, fy, j->dt);
        if (fabsf(VP_ORG_X - p->px) > VP_WDH * 0.55f || fabsf(VP_ORG_Y - p->py) > VP_WDH * 0.55f) {
            p->vx = 0.0f;
            p->vy = 0.0f;
        }
    }
    return NULL;
}

/* rs-src/nbody.rs:186-480.  Argument order is the reference's: theta first. */
void ora_step_barnes_hut(float theta, float dt, int32_t nthreads)
{
    if (theta == 0.0f) { /* rs-src/nbody.rs:197-200: brute force, and no velocity kill */
        ora_step_brute_force(dt);
        return;
    }
    ora_bh_build();
    if (nthreads <= 0) {
        /* rs-src/nbody.rs:424-428: the division by nthreads sits inside the closure mapped over
         * (0..nthreads); for nthreads <= 0 that range is empty, so the reference builds the tree, spawns
         * no thread and returns without moving any body. */
        g_tree_valid = 0;
        return;
    }
    /* rs-src/nbody.rs:424-428: range = n / nthreads; last thread takes the remainder */
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    BhJob *jobs = (BhJob *)malloc(sizeof(BhJob) * nthreads);
    int32_t range = g_n / nthreads;
    for (int32_t t = 0; t < nthreads; t++) {
        jobs[t].lo = range * t;
        jobs[t].hi = (t == nthreads - 1) ? g_n : range * (t + 1);
        jobs[t].theta = theta;
        jobs[t].dt = dt;
        if (nthreads == 1) bh_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, bh_worker, &jobs[t]);
    }
    if (nthreads > 1) for (int32_t t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
    g_tree_valid = 0; /* the tree described the OLD positions */
}

/* ---------------------------------------------------------------------------------------------
 * f64 helpers (not in the reference): the same 2-D pair law in double precision, used to measure
 * the rounding error of both the f32 reference semantics and the GPU kernels, and for the sampled
 * parity check at sizes the f32 restatement cannot finish.
 * ------------------------------------------------------------------------------------------- */
void ora_accel_f64_rows(const int32_t *rows, int32_t nrows, double *axy_out)
{
    for (int32_t r = 0; r < nrows; r++) {
        int32_t i = rows[r];
        double xi = g_p[i].px, yi = g_p[i].py, ax = 0.0, ay = 0.0;
        for (int32_t j = 0; j < g_n; j++) {
            if (j == i) continue;
            double dx = (double)g_p[j].px - xi, dy = (double)g_p[j].py - yi;
            double s = (double)g_p[j].m / (dx * dx + dy * dy + (double)EPS);
            ax += s * dx;
            ay += s * dy;
        }
        axy_out[2 * r] = ax;
        axy_out[2 * r + 1] = ay;
    }
}

/* ---------------------------------------------------------------------------------------------
 * 3-D extension checker (not in the reference, which is 2-D): f64 evaluation of the two pair laws the
 * nbx3_* kernels implement.  law 0: m_j d/(|d|^2+eps2)^(3/2) (Newtonian); law 1: m_j d/(|d|^2+eps2).
 * state7 = {px,py,pz,vx,vy,vz,m} per body.
 * ------------------------------------------------------------------------------------------- */
void ora3_accel_f64(const double *state7, int32_t n, int32_t law, double eps2, const int32_t *rows, int32_t nrows,
                    double *axyz_out)
{
    for (int32_t r = 0; r < nrows; r++) {
        const int32_t i = rows ? rows[r] : r;
        const double xi = state7[7 * i], yi = state7[7 * i + 1], zi = state7[7 * i + 2];
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int32_t j = 0; j < n; j++) {
            if (j == i) continue;
            const double dx = state7[7 * j] - xi, dy = state7[7 * j + 1] - yi, dz = state7[7 * j + 2] - zi;
            const double d2 = dx * dx + dy * dy + dz * dz + eps2;
            const double s = state7[7 * j + 6] / (law == 0 ? d2 * sqrt(d2) : d2);
            ax += s * dx; ay += s * dy; az += s * dz;
        }
        axyz_out[3 * r] = ax; axyz_out[3 * r + 1] = ay; axyz_out[3 * r + 2] = az;
    }
}
/* semi-implicit Euler (rs-src/nbody.rs:153-160) in f64 */
void ora3_step_f64(double *state7, int32_t n, int32_t law, double eps2, double dt)
{
    double *a = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    ora3_accel_f64(state7, n, law, eps2, NULL, n, a);
    for (int32_t i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) {
            state7[7 * i + 3 + k] += dt * a[3 * i + k];
            state7[7 * i + k] += dt * state7[7 * i + 3 + k];
        }
    }
    free(a);
}

/* ---------------------------------------------------------------------------------------------
 * Draw.  rs-src/nbody.rs:482-617.
 * ------------------------------------------------------------------------------------------- */

/* Rust `as i32` from f32: saturating, NaN -> 0 (defined behaviour since Rust 1.45; UB in 2016). */
static int32_t f32_as_i32(float v)
{
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}
static uint32_t f32_as_u32(float v)
{
    if (v != v) return 0;
    if (v >= 4294967296.0f) return UINT32_MAX;
    if (v <= 0.0f) return 0;
    return (uint32_t)v;
}

/* rs-src/nbody.rs:585-593 */
static uint32_t rgb_to_abgr32(uint8_t r8, uint8_t g8, uint8_t b8, float factor)
{
    uint32_t r = f32_as_u32((float)r8 * factor);
    uint32_t g = f32_as_u32((float)g8 * factor);
    uint32_t b = f32_as_u32((float)b8 * factor);
    return ((r > 255 ? 255 : r) << 0) | ((b > 255 ? 255 : b) << 16) | ((g > 255 ? 255 : g) << 8);
}
uint32_t ora_rgb_to_abgr32(uint8_t r, uint8_t g, uint8_t b, float factor) { return rgb_to_abgr32(r, g, b, factor); }

/* rs-src/nbody.rs:595-617 -- per-channel saturating add */
static uint32_t add_abgr32(uint32_t c1, uint32_t c2)
{
    uint32_t a1 = (c1 & 0xFF000000u) >> 24, b1 = (c1 & 0x00FF0000u) >> 16;
    uint32_t g1 = (c1 & 0x0000FF00u) >> 8, r1 = (c1 & 0x000000FFu);
    uint32_t a2 = (c2 & 0xFF000000u) >> 24, b2 = (c2 & 0x00FF0000u) >> 16;
    uint32_t g2 = (c2 & 0x0000FF00u) >> 8, r2 = (c2 & 0x000000FFu);
    uint32_t ar = a1 + a2 > 255 ? 255 : a1 + a2;
    uint32_t gr = g1 + g2 > 255 ? 255 : g1 + g2;
    uint32_t br = b1 + b2 > 255 ? 255 : b1 + b2;
    uint32_t rr = r1 + r2 > 255 ? 255 : r1 + r2;
    return (ar << 24) | (br << 16) | (gr << 8) | rr;
}
uint32_t ora_add_abgr32(uint32_t a, uint32_t b) { return add_abgr32(a, b); }

/* rs-src/nbody.rs:482-583 */
void ora_draw(int32_t w, int32_t h, uint32_t *fb)
{
    memset(fb, 0, (size_t)w * (size_t)h * sizeof(uint32_t));
    float aspect = (float)h / (float)w;
    float x1 = VP_ORG_X - VP_WDH / 2.0f;
    float y1 = (VP_ORG_Y - VP_WDH / 2.0f) * aspect;
    float x2 = VP_ORG_X + VP_WDH / 2.0f;
    float y2 = (VP_ORG_Y + VP_WDH / 2.0f) * aspect;
    float vpw = x2 - x1, vph = y2 - y1;
    float scalex = (1.0f / vpw) * (float)w;
    float scaley = (1.0f / vph) * (float)h;
    uint32_t col_body = rgb_to_abgr32(255, 215, 130, 0.3f);
    uint32_t col_tail = rgb_to_abgr32(255, 215, 130, 0.25f);
    static const int dir[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
    for (int32_t k = 0; k < g_n; k++) {
        const Particle *p = &g_p[k];
        float x = (p->px - x1) * scalex;
        float y = (p->py - y1) * scaley;
        for (int i = 0; i < 2; i++) {
            int32_t xo, yo;
            uint32_t col;
            if (i == 0) {
                xo = f32_as_i32(x);
                yo = f32_as_i32(y);
                col = col_body;
            } else {
                float angle = atan2f(p->vy, p->vx);
                int32_t octant = f32_as_i32(8.0f * angle / (2.0f * 3.14159265358979323846f) + 8.0f) % 8;
                if (octant < 0) octant += 8; /* unreachable for finite angle; Rust would index-panic */
                xo = f32_as_i32(x) - dir[octant][0];
                yo = f32_as_i32(y) - dir[octant][1];
                col = col_tail;
            }
            if (xo < 0 || xo >= w || yo < 0 || yo >= h) continue;
            int32_t idx = xo + yo * w;
            fb[idx] = add_abgr32(fb[idx], col);
        }
    }
    fb[w / 2 + 0 + (h / 2 + 0) * w] = 0x00FF00FFu;
    fb[w / 2 + 1 + (h / 2 + 0) * w] = 0x00FF00FFu;
    fb[w / 2 + 0 + (h / 2 + 1) * w] = 0x00FF00FFu;
    fb[w / 2 - 1 + (h / 2 + 0) * w] = 0x00FF00FFu;
    fb[w / 2 + 0 + (h / 2 - 1) * w] = 0x00FF00FFu;
}
