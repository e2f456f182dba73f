/*
 * results_oracle.c -- CPU restatement of the host post-processing that follows the score loop.
 * TEST INFRASTRUCTURE ONLY (see myrio_oracle.h).
 *
 * Follows, line by line (paths relative to the reference root, Myrio v0.3.0):
 *   myrio-core/src/tax/results.rs:96-143    dive_recursive: per-branch nb_leaves / sum / max over ALL leaves
 *   myrio-core/src/tax/results.rs:257-307   trim(n): the n best direct leaves of every branch, renumbered
 *   myrio-core/src/tax/results.rs:310-381   cut(x): global top-x of the (trimmed) leaves, tree rebuilt
 *   myrio-core/src/tax/results.rs:160-248   softmax pooling per rank, confidence, force_keep
 *   myrio-core/src/tax/core.rs:180-208      gather_branches_at_rank (height = root_rank - depth)
 *   myrio-cli/src/app/process/run.rs:405-466  in-database confidence over the genes of a run
 *
 * The tree is built with nested structs like the reference (the library works on flat arrays); f32 arithmetic in
 * the reference's order, expf / powf from libm as Rust's f32::exp / f32::powf call them.
 *
 * PARITY UNPINNED where the reference itself is not deterministic or delegates to itertools 0.14.0:
 *   - k_largest_by_key ties (results.rs:221-225, 286-293): defined here as (score desc, original order asc);
 *   - the children order of a branch comes from a HashMap in the reference (store.rs:205-236): here it is the
 *     order of the input node list.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "myrio_oracle.h"

typedef struct rnode {
    int is_leaf;
    uint64_t src;        /* index of the node in the input list (carries the name) */
    uint64_t payload_id; /* leaf */
    struct rnode **children;
    size_t n_children;
    /* BRes (results.rs:23-30) */
    uint64_t nb_leaves;
    float mean, max, pool_score;
    int conf_state; /* 0 = None, 1 = Some(None), 2 = Some(Some(conf)) */
    float conf;
    int force_keep;
} rnode;

static rnode *node_new(int is_leaf, uint64_t src) {
    rnode *n = (rnode *)calloc(1, sizeof(rnode));
    n->is_leaf = is_leaf;
    n->src = src;
    return n;
}
static void node_push(rnode *b, rnode *c) {
    b->children = (rnode **)realloc(b->children, (b->n_children + 1) * sizeof(rnode *));
    b->children[b->n_children++] = c;
}
static void node_free(rnode *n) {
    if (!n) return;
    for (size_t i = 0; i < n->n_children; ++i) node_free(n->children[i]);
    free(n->children);
    free(n);
}

/* results.rs:96-143 */
typedef struct {
    uint64_t nb;
    float sum, max;
} dive_t;

static rnode *dive(const uint8_t *kind, const uint64_t *arg, uint64_t *pos, const float *simscores, dive_t *out) {
    uint64_t i = (*pos)++;
    if (kind[i] == 1) { /* Node::Leaf */
        rnode *l = node_new(1, i);
        l->payload_id = arg[i];
        float score = simscores[arg[i]];
        out->nb = 1;
        out->sum = score;
        out->max = score;
        return l;
    }
    rnode *b = node_new(0, i);
    uint64_t nb_leaves = 0;
    float sum = 0.0f, max = 0.0f;
    for (uint64_t c = 0; c < arg[i]; ++c) {
        dive_t d;
        node_push(b, dive(kind, arg, pos, simscores, &d));
        nb_leaves += d.nb;
        sum = sum + d.sum;
        max = fmaxf(max, d.max);
    }
    b->nb_leaves = nb_leaves;
    b->mean = sum / (float)nb_leaves;
    b->max = max;
    out->nb = nb_leaves;
    out->sum = sum;
    out->max = max;
    return b;
}

typedef struct {
    float *score;
    uint64_t *orig; /* payload id in the tree the results were computed from */
    size_t n, cap;
} payloads_t;
static uint64_t payloads_push(payloads_t *p, float s, uint64_t orig) {
    if (p->n == p->cap) {
        p->cap = p->cap ? p->cap * 2 : 64;
        p->score = (float *)realloc(p->score, p->cap * sizeof(float));
        p->orig = (uint64_t *)realloc(p->orig, p->cap * sizeof(uint64_t));
    }
    p->score[p->n] = s;
    p->orig[p->n] = orig;
    return p->n++;
}

/* results.rs:257-307 trim */
static void trim_rec(rnode *node, payloads_t *newp, const float *simscores, const uint64_t *orig_of, size_t max) {
    if (node->is_leaf) return;
    rnode **children = node->children;
    size_t n = node->n_children;
    rnode **new_children = NULL, **leaves = NULL;
    size_t n_new = 0, n_leaves = 0;
    new_children = (rnode **)malloc((n + 1) * sizeof(rnode *));
    leaves = (rnode **)malloc((n + 1) * sizeof(rnode *));
    for (size_t i = 0; i < n; ++i) {
        if (children[i]->is_leaf)
            leaves[n_leaves++] = children[i];
        else
            new_children[n_new++] = children[i];
    }
    for (size_t i = 0; i < n_new; ++i) trim_rec(new_children[i], newp, simscores, orig_of, max);
    /* k_largest_by_key(max, score): descending; ties keep the original order (see header) */
    uint8_t *used = (uint8_t *)calloc(n_leaves + 1, 1);
    size_t take = n_leaves < max ? n_leaves : max;
    for (size_t t = 0; t < take; ++t) {
        size_t best = (size_t)-1;
        for (size_t j = 0; j < n_leaves; ++j) {
            if (used[j]) continue;
            if (best == (size_t)-1 || simscores[leaves[j]->payload_id] > simscores[leaves[best]->payload_id]) best = j;
        }
        used[best] = 1;
        rnode *l = leaves[best];
        float s = simscores[l->payload_id];
        l->payload_id = payloads_push(newp, s, orig_of ? orig_of[l->payload_id] : l->payload_id);
        new_children[n_new++] = l;
    }
    for (size_t j = 0; j < n_leaves; ++j)
        if (!used[j]) node_free(leaves[j]);
    free(used);
    free(leaves);
    free(children);
    node->children = new_children;
    node->n_children = n_new;
}

/* Node::gather_leaves_and_branches: does a leaf of `best` / a force_keep branch sit at or below `n`? */
static int has_best_or_forced(const rnode *n, const uint8_t *in_best) {
    if (n->is_leaf) return in_best[n->payload_id];
    if (n->force_keep) return 1;
    for (size_t i = 0; i < n->n_children; ++i)
        if (has_best_or_forced(n->children[i], in_best)) return 1;
    return 0;
}

/* results.rs:329-366 */
static void cut_rec(const rnode *node, rnode *above, payloads_t *newp, const uint8_t *in_best, const payloads_t *old) {
    if (node->is_leaf) {
        if (in_best[node->payload_id]) {
            rnode *l = node_new(1, node->src);
            l->payload_id = payloads_push(newp, old->score[node->payload_id], old->orig[node->payload_id]);
            node_push(above, l);
        }
        return;
    }
    int add_self = node->force_keep;
    rnode *cur = node_new(0, node->src);
    for (size_t i = 0; i < node->n_children; ++i) {
        const rnode *sub = node->children[i];
        if (has_best_or_forced(sub, in_best)) {
            add_self = 1;
            cut_rec(sub, cur, newp, in_best, old);
        }
    }
    if (add_self) {
        cur->nb_leaves = node->nb_leaves;
        cur->mean = node->mean;
        cur->max = node->max;
        cur->pool_score = node->pool_score;
        cur->conf_state = node->conf_state;
        cur->conf = node->conf;
        cur->force_keep = node->force_keep;
        node_push(above, cur);
    } else {
        node_free(cur);
    }
}

static void gather_leaves(const rnode *n, const rnode ***out, size_t *cnt, size_t *cap) {
    if (n->is_leaf) {
        if (*cnt == *cap) {
            *cap = *cap ? *cap * 2 : 64;
            *out = (const rnode **)realloc((void *)*out, *cap * sizeof(rnode *));
        }
        (*out)[(*cnt)++] = n;
        return;
    }
    for (size_t i = 0; i < n->n_children; ++i) gather_leaves(n->children[i], out, cnt, cap);
}

/* results.rs:310-381 cut: returns the new root (NULL if nothing is kept) */
static rnode *cut_tree(const rnode *root, const payloads_t *old, size_t nb_best, payloads_t *newp) {
    const rnode **leaves = NULL;
    size_t n = 0, cap = 0;
    gather_leaves(root, &leaves, &n, &cap);
    /* sorted_by_key(-score): stable, so equal scores stay in gather order; take(nb_best) */
    size_t *order = (size_t *)malloc((n + 1) * sizeof(size_t));
    for (size_t i = 0; i < n; ++i) order[i] = i;
    for (size_t i = 1; i < n; ++i) { /* insertion sort = stable */
        size_t v = order[i];
        float sv = old->score[leaves[v]->payload_id];
        size_t j = i;
        while (j > 0 && old->score[leaves[order[j - 1]]->payload_id] < sv) {
            order[j] = order[j - 1];
            --j;
        }
        order[j] = v;
    }
    uint8_t *in_best = (uint8_t *)calloc(old->n + 1, 1);
    for (size_t i = 0; i < n && i < nb_best; ++i) in_best[leaves[order[i]]->payload_id] = 1;
    rnode wrap;
    memset(&wrap, 0, sizeof(wrap));
    cut_rec(root, &wrap, newp, in_best, old);
    rnode *res = wrap.n_children ? wrap.children[0] : NULL; /* extract_singlevec */
    free(wrap.children);
    free(in_best);
    free(order);
    free((void *)leaves);
    return res;
}

/* results.rs:160-173 */
static float softmax_pooling(const float *scores, size_t n, float lambda) {
    float numerator = 0.0f, denominator = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float s = scores[i];
        numerator = numerator + s * expf(lambda * s);
        denominator = denominator + expf(lambda * s);
    }
    return numerator / denominator;
}

/* core.rs:180-208 */
static void branches_at(rnode *n, size_t cur_h, size_t want_h, rnode ***out, size_t *cnt, size_t *cap) {
    if (n->is_leaf) return;
    if (cur_h == want_h) {
        if (*cnt == *cap) {
            *cap = *cap ? *cap * 2 : 16;
            *out = (rnode **)realloc(*out, *cap * sizeof(rnode *));
        }
        (*out)[(*cnt)++] = n;
        return;
    }
    for (size_t i = 0; i < n->n_children; ++i) branches_at(n->children[i], cur_h - 1, want_h, out, cnt, cap);
}

/* results.rs:175-248; returns 0, or -40 when a pool score is not finite (SimScore::try_new(..).unwrap() panics) */
static int pool_and_confidence(rnode *root, uint32_t root_rank, const payloads_t *pl, const myo_results_params *p) {
    for (uint32_t rank = 2; rank <= root_rank; ++rank) { /* Rank::collect_range_inclusive(Genus, root_rank) */
        rnode **v = NULL;
        size_t n = 0, cap = 0;
        branches_at(root, root_rank, rank, &v, &n, &cap);
        for (size_t b = 0; b < n; ++b) {
            rnode *br = v[b];
            float *sims = (float *)malloc((br->n_children + 1) * sizeof(float));
            float *pools = (float *)malloc((br->n_children + 2) * sizeof(float));
            size_t ns = 0, np = 0;
            for (size_t c = 0; c < br->n_children; ++c) {
                rnode *s = br->children[c];
                if (s->is_leaf)
                    sims[ns++] = pl->score[s->payload_id];
                else
                    pools[np++] = s->pool_score;
            }
            if (ns) {
                float nf = (float)ns;
                float penalty = nf / (nf + p->mu);
                pools[np++] = softmax_pooling(sims, ns, p->lambda_leaf) * penalty;
            }
            br->pool_score = softmax_pooling(pools, np, p->lambda_branch);
            free(sims);
            free(pools);
        }
        if (n == 0) {
            free(v);
            continue;
        }
        if (n == 1) {
            if (rank != root_rank) v[0]->conf_state = 1; /* Some(None) */
            free(v);
            continue;
        }
        float nf = (float)n;
        float sum = 0.0f;
        for (size_t b = 0; b < n; ++b) sum = sum + v[b]->pool_score;
        float mean = sum / nf;
        float sq = 0.0f;
        for (size_t b = 0; b < n; ++b) {
            float d = v[b]->pool_score - mean;
            sq = sq + d * d;
        }
        float sd = fmaxf(sqrtf((1.0f / nf) * sq), 1E-3f);
        /* k_largest_by_key(2, pool_score): ties -> the earlier branch first */
        size_t first = (size_t)-1, second = (size_t)-1;
        for (size_t b = 0; b < n; ++b) {
            if (!isfinite(v[b]->pool_score)) {
                free(v);
                return -40;
            }
            if (first == (size_t)-1 || v[b]->pool_score > v[first]->pool_score) {
                second = first;
                first = b;
            } else if (second == (size_t)-1 || v[b]->pool_score > v[second]->pool_score) {
                second = b;
            }
        }
        float x = (v[first]->pool_score - v[second]->pool_score) / sd;
        float conf = powf(v[first]->pool_score, p->epsilon) * powf(x, p->gamma) / (p->delta + powf(x, p->gamma));
        v[first]->conf_state = 2;
        v[first]->conf = conf;
        v[first]->force_keep = 1;
        if (v[second]->pool_score > 0.2f && conf < 0.3f) v[second]->force_keep = 1;
        free(v);
    }
    return 0;
}

/* ---- flat output ---- */

struct myo_results {
    uint64_t n_nodes, cap;
    uint8_t *kind;
    uint64_t *arg, *src, *nb_leaves;
    float *mean, *max, *pool, *conf;
    uint8_t *conf_state, *force_keep;
    payloads_t pl;
    rnode *root; /* kept for myo_results_cut */
    uint32_t root_rank;
};

static void flat_push(myo_results *r, const rnode *n) {
    if (r->n_nodes == r->cap) {
        r->cap = r->cap ? r->cap * 2 : 64;
        r->kind = (uint8_t *)realloc(r->kind, r->cap);
        r->arg = (uint64_t *)realloc(r->arg, r->cap * 8);
        r->src = (uint64_t *)realloc(r->src, r->cap * 8);
        r->nb_leaves = (uint64_t *)realloc(r->nb_leaves, r->cap * 8);
        r->mean = (float *)realloc(r->mean, r->cap * 4);
        r->max = (float *)realloc(r->max, r->cap * 4);
        r->pool = (float *)realloc(r->pool, r->cap * 4);
        r->conf = (float *)realloc(r->conf, r->cap * 4);
        r->conf_state = (uint8_t *)realloc(r->conf_state, r->cap);
        r->force_keep = (uint8_t *)realloc(r->force_keep, r->cap);
    }
    uint64_t i = r->n_nodes++;
    r->kind[i] = (uint8_t)n->is_leaf;
    r->arg[i] = n->is_leaf ? n->payload_id : n->n_children;
    r->src[i] = n->src;
    r->nb_leaves[i] = n->nb_leaves;
    r->mean[i] = n->mean;
    r->max[i] = n->max;
    r->pool[i] = n->pool_score;
    r->conf[i] = n->conf;
    r->conf_state[i] = (uint8_t)n->conf_state;
    r->force_keep[i] = (uint8_t)n->force_keep;
    for (size_t c = 0; c < n->n_children; ++c) flat_push(r, n->children[c]);
}

static myo_results *results_wrap(rnode *root, payloads_t pl, uint32_t root_rank) {
    myo_results *r = (myo_results *)calloc(1, sizeof(*r));
    r->root = root;
    r->pl = pl;
    r->root_rank = root_rank;
    if (root) flat_push(r, root);
    return r;
}

/* TaxTreeResults::from_compute_tree after the score loop (results.rs:96-252) */
int myo_results_from_scores(uint32_t root_rank, uint64_t n_nodes, const uint8_t *kind, const uint64_t *arg,
                            const float *simscores, uint64_t n_payloads, const myo_results_params *p,
                            myo_results **out) {
    (void)n_nodes;
    (void)n_payloads;
    uint64_t pos = 0;
    dive_t d;
    rnode *root = dive(kind, arg, &pos, simscores, &d);
    payloads_t trimmed = {0};
    trim_rec(root, &trimmed, simscores, NULL, (size_t)p->max_amount_of_leaves_per_branch);
    payloads_t cutp = {0};
    rnode *cut = cut_tree(root, &trimmed, (size_t)p->nb_best_analysis, &cutp);
    node_free(root);
    free(trimmed.score);
    free(trimmed.orig);
    if (!cut) { /* extract_singlevec on an empty vector panics */
        free(cutp.score);
        free(cutp.orig);
        return -41;
    }
    int rc = pool_and_confidence(cut, root_rank, &cutp, p);
    if (rc) {
        node_free(cut);
        free(cutp.score);
        free(cutp.orig);
        return rc;
    }
    *out = results_wrap(cut, cutp, root_rank);
    return 0;
}

/* TaxTreeResults::cut on an analysed tree (run.rs:392: cut(nb_best_display)); force_keep is honoured */
int myo_results_cut(const myo_results *r, uint64_t nb_best, myo_results **out) {
    payloads_t np = {0};
    rnode *cut = cut_tree(r->root, &r->pl, (size_t)nb_best, &np);
    if (!cut) {
        free(np.score);
        free(np.orig);
        return -41;
    }
    *out = results_wrap(cut, np, r->root_rank);
    return 0;
}

void myo_results_view(const myo_results *r, myo_results_view_t *v) {
    v->n_nodes = r->n_nodes;
    v->n_payloads = r->pl.n;
    v->kind = r->kind;
    v->arg = r->arg;
    v->src = r->src;
    v->nb_leaves = r->nb_leaves;
    v->mean = r->mean;
    v->max = r->max;
    v->pool_score = r->pool;
    v->conf = r->conf;
    v->conf_state = r->conf_state;
    v->force_keep = r->force_keep;
    v->payload_score = r->pl.score;
    v->payload_orig = r->pl.orig;
}

void myo_results_free(myo_results *r) {
    if (!r) return;
    node_free(r->root);
    free(r->kind);
    free(r->arg);
    free(r->src);
    free(r->nb_leaves);
    free(r->mean);
    free(r->max);
    free(r->pool);
    free(r->conf);
    free(r->conf_state);
    free(r->force_keep);
    free(r->pl.score);
    free(r->pl.orig);
    free(r);
}

/* run.rs:405-466.  Per gene g the genus branches of its results tree: name ids (equal ids = equal names),
 * pool scores, confidence state / value, genus_off[g] .. genus_off[g+1].  Returns the in-database confidence. */
float myo_in_database_confidence(uint64_t n_genes, const uint64_t *genus_off, const uint64_t *name_id,
                                 const float *pool_score, const uint8_t *conf_state, const float *conf) {
    enum { TAKE = 8 };
    /* sorted_unstable_by_key(-score) per gene: descending pool score (ties: input order) */
    uint64_t total = genus_off[n_genes];
    uint64_t *ord = (uint64_t *)malloc((total + 1) * sizeof(uint64_t));
    for (uint64_t g = 0; g < n_genes; ++g) {
        uint64_t b = genus_off[g], e = genus_off[g + 1];
        for (uint64_t i = b; i < e; ++i) ord[i] = i;
        for (uint64_t i = b + 1; i < e; ++i) {
            uint64_t v = ord[i], j = i;
            while (j > b && pool_score[ord[j - 1]] < pool_score[v]) {
                ord[j] = ord[j - 1];
                --j;
            }
            ord[j] = v;
        }
    }
    float mismatch = 0.0f;
    float *top_conf = (float *)malloc((n_genes + 1) * sizeof(float));
    for (uint64_t i = 0; i < n_genes; ++i) {
        uint64_t top = ord[genus_off[i]];
        float c = conf_state[top] == 2 ? conf[top] : 0.5f; /* unwrap().unwrap_or_else(0.5) */
        top_conf[i] = c;
        for (uint64_t o = i + 1; o < n_genes; ++o) {
            uint64_t b = genus_off[o], e = genus_off[o + 1];
            uint64_t posn = TAKE;
            for (uint64_t t = 0; t < TAKE && b + t < e; ++t)
                if (name_id[ord[b + t]] == name_id[top]) {
                    posn = t;
                    break;
                }
            float penalty = (float)posn / (float)TAKE;
            mismatch = mismatch + penalty * fmaxf(powf(c, 0.7f), 0.12f);
        }
    }
    mismatch = mismatch / (float)n_genes;
    float csum = 0.0f;
    for (uint64_t i = 0; i < n_genes; ++i) csum = csum + top_conf[i];
    float lambda = 0.5f + 2.0f * csum;
    float conf_score = 1.0f - softmax_pooling(top_conf, n_genes, lambda);
    float x = mismatch, y = powf(conf_score, 0.8f);
    float unknown = 1.0f / (1.0f + expf(-8.0f * ((x + y) / 2.0f - 0.45f))); /* powi(-1) */
    free(ord);
    free(top_conf);
    return 1.0f - unknown;
}
