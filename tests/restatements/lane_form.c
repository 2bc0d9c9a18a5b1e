/* lane_form.c -- CPU restatement of the FORMULATION the lane-per-block ClusterFit kernels use
 * (texpresso_b200/csrc/txp_cluster_lane.cuh), checked against the oracle's literal loop nest.  Test infrastructure only.
 *
 * What differs from the reference's text (cluster.rs:303-381) and therefore needs an argument that the bits are the same:
 *   1. the (i, j)-only halves of alphax_sum / betax_sum, A = part1*(2/3,4/9) + part0 and B = part1*(1/3,1/9), are
 *      computed once per (i, j);  per candidate alphax = part2*(1/3,1/9) + A and betax = B + (part2*(2/3,4/9) + part3);
 *   2. part0 / part1 / part2 are advanced unconditionally, with a zero guard entry at points_weights[count]
 *      (the reference guards the additions with `if k < count`);
 *   3. the j == 0 row starts from part2 = points_weights[0], k = 1, exactly as the reference -- the kernel keeps that;
 *   4. the winner is carried as (error, key = i<<10 | j<<5 | k) with a strict `<`, and its endpoints are recomputed
 *      afterwards from freshly accumulated part sums (left to right from zero).
 * The functions below replace cluster_compress4 by that formulation inside the oracle's own ClusterFit state machine
 * (orderings, iteration rule, remap, write4 are the oracle's) and count the blocks whose 8 output bytes differ.
 * Build: gcc -O2 -ffp-contract=off (tests/test_lane_formulation.py). */
#include "../../oracle/txp_oracle.c"

static v4 range_sum(const v4 *pw, int a, int b) {
    v4 acc = {0, 0, 0, 0};
    for (int m = a; m < b; ++m) acc = v4add(acc, pw[m]);
    return acc;
}

static void cluster_compress4_lane(clusterfit *f) {
    const int count = f->set->count;
    const v4 zero = {0, 0, 0, 0};
    const v4 c13 = {1.0f / 3.0f, 1.0f / 3.0f, 1.0f / 3.0f, 1.0f / 9.0f};
    const v4 c23 = {2.0f / 3.0f, 2.0f / 3.0f, 2.0f / 3.0f, 4.0f / 9.0f};
    const float twoninths = 2.0f / 9.0f;
    v3 best_start = {0, 0, 0}, best_end = {0, 0, 0};
    float run_best = f->best_error;
    int best_iteration = 0;
    uint32_t best_key = 0xFFFFFFFFu;
    uint8_t best_order[16] = {0};
    v3 axis = f->principle;

    for (int it = 0; it < f->num_iterations; ++it) {
        if (!construct_ordering(f, axis, it)) break;
        v4 pw[17];
        for (int m = 0; m < count; ++m) pw[m] = f->pw[m];
        pw[count] = zero;                                              /* guard entry */
        const v4 xsum = range_sum(pw, 0, count);
        float err_it = run_best;
        uint32_t key_it = 0xFFFFFFFFu;
        v4 part0 = zero;
        for (int i = 0; i < count; ++i) {
            v4 part1 = zero;
            for (int j = i; j <= count; ++j) {
                const v4 A = v4add(v4mul(part1, c23), part0), B = v4mul(part1, c13);
                v4 part2 = zero;
                int k0 = j;
                if (j == 0) { part2 = pw[0]; k0 = 1; }
                for (int k = k0; k <= count; ++k) {
                    const v4 part3 = v4sub(v4sub(v4sub(xsum, part2), part1), part0);
                    const v4 alphax = v4add(v4mul(part2, c13), A);
                    const v4 betax = v4add(B, v4add(v4mul(part2, c23), part3));
                    const float alphabeta = twoninths * (part1.w + part2.w);
                    const lsq r = solve(alphax, betax, alphabeta, f->mw);
                    if (r.error < err_it) { err_it = r.error; key_it = ((uint32_t)i << 10) | ((uint32_t)j << 5) | (uint32_t)k; }
                    part2 = v4add(part2, pw[k]);                       /* unconditional: pw[count] is zero */
                }
                part1 = v4add(part1, pw[j]);
            }
            part0 = v4add(part0, pw[i]);
        }
        if (key_it != 0xFFFFFFFFu) {                                   /* strictly better than everything before */
            const int bi = (int)(key_it >> 10), bj = (int)((key_it >> 5) & 31u), bk = (int)(key_it & 31u);
            const v4 p0 = range_sum(pw, 0, bi), p1 = range_sum(pw, bi, bj), p2 = range_sum(pw, bj, bk);
            const v4 p3 = v4sub(v4sub(v4sub(xsum, p2), p1), p0);
            const v4 alphax = v4add(v4mul(p2, c13), v4add(v4mul(p1, c23), p0));
            const v4 betax = v4add(v4mul(p1, c13), v4add(v4mul(p2, c23), p3));
            const lsq r = solve(alphax, betax, twoninths * (p1.w + p2.w), f->mw);
            best_start.x = r.ax; best_start.y = r.ay; best_start.z = r.az;
            best_end.x = r.bx; best_end.y = r.by; best_end.z = r.bz;
            run_best = err_it; best_key = key_it; best_iteration = it;
            memcpy(best_order, f->order[it], 16);
        }
        if (best_iteration != it) break;
        axis.x = best_end.x - best_start.x; axis.y = best_end.y - best_start.y; axis.z = best_end.z - best_start.z;
    }

    if (run_best < f->best_error) {
        const int bi = (int)(best_key >> 10), bj = (int)((best_key >> 5) & 31u), bk = (int)(best_key & 31u);
        uint8_t unordered[16], best_indices[16];
        memset(unordered, 0, 16);
        for (int m = 0; m < count; ++m)                                /* one ascending pass, later writes win (kernel form) */
            unordered[best_order[m]] = m < bi ? 0 : (m < bj ? 2 : (m < bk ? 3 : 1));
        remap_indices(f->set, unordered, best_indices);
        write4(best_start, best_end, best_indices, f->best_compressed);
        f->best_error = run_best;
    }
}

/* number of blocks (of n) whose colour bytes differ between the oracle and the lane formulation; blocks with fewer than
 * two points do not reach ClusterFit and are skipped.  *searched receives how many blocks were compared. */
TXO_API size_t txl_compare(int fmt, const uint8_t *blocks, const uint32_t *masks, size_t n, const txo_params *p, size_t *searched) {
    size_t bad = 0, done = 0;
    for (size_t b = 0; b < n; ++b) {
        colourset set;
        colourset_new(&set, blocks + 64 * b, masks[b], fmt, p->weigh_colour_by_alpha != 0);
        if (set.count < 2) continue;
        uint8_t want[8], got[8];
        cluster_compress(&set, fmt, p->weights, p->algorithm == TXO_ITERATIVE, want, NULL);
        clusterfit f;
        memset(&f, 0, sizeof f);
        f.set = &set; f.fmt = fmt; f.mw[0] = p->weights[0]; f.mw[1] = p->weights[1]; f.mw[2] = p->weights[2];
        f.num_iterations = p->algorithm == TXO_ITERATIVE ? 8 : 1;
        f.best_error = FLT_MAX;
        float cov[6];
        weighted_covariance(set.points, set.weights, set.count, cov);
        f.principle = principle_component(cov);
        if (fmt == TXO_BC1) {
            cluster_compress3(&f);
            if (!set.transparent) cluster_compress4_lane(&f);
        } else {
            cluster_compress4_lane(&f);
        }
        memcpy(got, f.best_compressed, 8);
        ++done;
        if (memcmp(want, got, 8) != 0) ++bad;
    }
    if (searched) *searched = done;
    return bad;
}
