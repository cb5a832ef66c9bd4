/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the two moolib hot paths.
 *
 * Nothing in the product (moolib_b200/) may include, link or call this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs do, and only as the checker.
 *
 * Each function cites the reference code it restates (paths relative to the reference tree @ 06e7a3e).  The
 * restatement is pinned against the compiled reference itself (oracle/_ref, built by oracle/build_ref.sh) by
 * tests/test_oracle_pinning.py and against the fixtures in tests/golden/ that were generated from it.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, IEEE fp32 adds/multiplies).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------------
 * HP-B: batch gather / stack / cat.  Pure byte movement.
 * ------------------------------------------------------------------------------------------------------------ */

/* src/env.h:248-263 Env::fillBatch: dst = slab + itemsize*elements*batchIndex; memcpy(dst, src, itemsize*elements) */
void oracle_fill_batch(uint8_t* slab, size_t itemsize, size_t elements, size_t batch_index, const void* src) {
  memcpy(slab + itemsize * elements * batch_index, src, itemsize * elements);
}

/* src/moolib.cc:676 / :751  target.select(dim, slot).copy_(src)
 * target is contiguous [outer, size, inner_bytes], src contiguous [outer, inner_bytes]. */
void oracle_stack_slot(uint8_t* dst, size_t outer, size_t size, size_t slot, size_t inner_bytes, const uint8_t* src) {
  for (size_t o = 0; o < outer; ++o) memcpy(dst + (o * size + slot) * inner_bytes, src + o * inner_bytes, inner_bytes);
}

/* src/moolib.cc:665-668 / :745-748  dst.narrow(dim, dst_off, n).copy_(src.narrow(dim, src_off, n))
 * dst contiguous [outer, dst_dim, inner_bytes], src contiguous [outer, src_dim, inner_bytes]. */
void oracle_cat_narrow(uint8_t* dst, const uint8_t* src, size_t outer, size_t dst_dim, size_t dst_off, size_t src_dim,
                       size_t src_off, size_t n, size_t inner_bytes) {
  for (size_t o = 0; o < outer; ++o)
    memcpy(dst + (o * dst_dim + dst_off) * inner_bytes, src + (o * src_dim + src_off) * inner_bytes, n * inner_bytes);
}

/* the generic pitched copy both of the above reduce to (the unit of work of mb_copy2d_batch) */
void oracle_copy2d(const uint8_t* src, uint8_t* dst, size_t row_bytes, size_t rows, ptrdiff_t src_pitch,
                   ptrdiff_t dst_pitch) {
  for (size_t r = 0; r < rows; ++r) memcpy(dst + (ptrdiff_t)r * dst_pitch, src + (ptrdiff_t)r * src_pitch, row_bytes);
}

/* src/env.cc:340-345  envInputs[i].action.store(action.load() + 1 + acc[i])  (uint32 wrap-around arithmetic) */
void oracle_scatter_actions(uint32_t* counters, size_t stride_u32, const int64_t* actions, size_t n) {
  for (size_t i = 0; i < n; ++i) counters[i * stride_u32] = counters[i * stride_u32] + 1u + (uint32_t)actions[i];
}

/* ------------------------------------------------------------------------------------------------------------
 * HP-A: gradient staging, tree allreduce, scale.
 * ------------------------------------------------------------------------------------------------------------ */

typedef struct oracle_hdr {
  uint64_t num_gradients, num_skipped, batch_size, has_grads;
} oracle_hdr;

/* src/accumulator.cc:941-980: first contribution copies (targetGradients[i].copy_(grad)), later ones accumulate
 * (targetGradients[i].add_(addGrads[i]));  :410-418 actuallyZeroGradients zeroes the sources afterwards.
 * Layout of `staging`: tensor i at float offset sum_{j<i} roundup4(numel_j) (the product's flat layout; padding = 0). */
void oracle_stage(float* staging, float* const* grads, const uint64_t* numel, int ntensors, int accumulate,
                  int zero_src) {
  uint64_t off = 0;
  for (int t = 0; t < ntensors; ++t) {
    const uint64_t n = numel[t], padded = (n + 3) & ~(uint64_t)3;
    for (uint64_t i = 0; i < n; ++i) {
      if (accumulate) staging[off + i] = staging[off + i] + grads[t][i];
      else staging[off + i] = grads[t][i];
      if (zero_src) grads[t][i] = 0.0f;
    }
    if (!accumulate)
      for (uint64_t i = n; i < padded; ++i) staging[off + i] = 0.0f;
    off += padded;
  }
}

/* One node's payload in the tree: gradient vector (NULL = the peer skipped, i.e. an EMPTY gradient list) + counters. */
typedef struct node_val {
  float* g; /* owned scratch or NULL */
  oracle_hdr h;
} node_val;

/* src/group.h:201-212 AccumulatorReductionType::add */
static void payload_add(node_val* a, node_val* n, size_t numel) {
  if ((n->g != NULL) == (a->g != NULL)) {
    if (a->g)
      for (size_t i = 0; i < numel; ++i) a->g[i] = a->g[i] + n->g[i]; /* gradients[i] += n.gradients[i] */
  } else if (n->g != NULL) { /* n.gradients.size() > gradients.size(): std::swap */
    float* t = a->g;
    a->g = n->g;
    n->g = t;
  }
  a->h.num_gradients += n->h.num_gradients;
  a->h.num_skipped += n->h.num_skipped;
  a->h.batch_size += n->h.batch_size;
}

/* Binary-tree reduce of src/group.h:570-629 + 738-768: peer p>=1 receives from children 2p and 2p+1, the root 0 from
 * its single child 1; `local += remote` in ARRIVAL order, which the reference leaves to the network.  `order` is a
 * bitmask that picks one legal arrival order: bit p clear = node p adds its lower-index child first, set = the
 * higher-index child first. */
static void tree_reduce(node_val* v, int npeers, int p, size_t numel, int order) {
  int kids[2], nk = 0;
  if (p == 0) {
    if (npeers > 1) kids[nk++] = 1;
  } else {
    if (2 * p < npeers) kids[nk++] = 2 * p;
    if (2 * p + 1 < npeers) kids[nk++] = 2 * p + 1;
  }
  if (nk == 2 && ((order >> p) & 1)) {
    int t = kids[0];
    kids[0] = kids[1];
    kids[1] = t;
  }
  for (int k = 0; k < nk; ++k) {
    tree_reduce(v, npeers, kids[k], numel, order);
    payload_add(&v[p], &v[kids[k]], numel);
  }
}

/* Full reference semantics of one Accumulator gradient round:
 *   in[r]   gradient vector of peer r (flat, numel floats) or NULL if the peer called skip_gradients()
 *   hdr[r]  that peer's {numGradients,numSkipped,batchSize}
 *   out     what EVERY peer's .grad holds afterwards (src/group.h:553-568 share = broadcast of the root's result;
 *           src/accumulator.cc:425-452 setGradients: zero if the result has no gradients, else
 *           grad.copy_(sum); grad.mul_(1.0f / numGradients) -- fp32 multiply by the fp32 reciprocal)
 * scale == 0 gives the plain sum (GroupWrapper::allReduce with ReduceSum, src/group.h:243-247).
 * Returns 0, or -1 on allocation failure. */
int oracle_allreduce_tree(const float* const* in, const oracle_hdr* hdr, int npeers, size_t numel, int order, int scale,
                          float* out, oracle_hdr* out_hdr) {
  node_val* v = (node_val*)calloc((size_t)npeers, sizeof(node_val));
  if (!v) return -1;
  for (int r = 0; r < npeers; ++r) {
    v[r].h = hdr[r];
    if (in[r]) {
      v[r].g = (float*)malloc(numel * sizeof(float) + 1);
      if (!v[r].g) return -1;
      memcpy(v[r].g, in[r], numel * sizeof(float));
    }
  }
  tree_reduce(v, npeers, 0, numel, order);
  oracle_hdr tot = v[0].h;
  tot.has_grads = 0;
  for (int r = 0; r < npeers; ++r) tot.has_grads += in[r] != NULL;
  if (!v[0].g) {
    for (size_t i = 0; i < numel; ++i) out[i] = 0.0f; /* accumulator.cc:426-428 */
  } else if (scale && tot.num_gradients) {
    const float s = 1.0f / (float)tot.num_gradients;
    for (size_t i = 0; i < numel; ++i) out[i] = v[0].g[i] * s;
  } else {
    memcpy(out, v[0].g, numel * sizeof(float));
  }
  if (out_hdr) *out_hdr = tot;
  for (int r = 0; r < npeers; ++r) free(v[r].g);
  free(v);
  return 0;
}

/* The product's summation order (ascending rank over the peers that have gradients), same scale rule.  The product
 * must match THIS bit for bit; it must match oracle_allreduce_tree within the tolerance DESIGN.md states. */
int oracle_allreduce_rankorder(const float* const* in, const oracle_hdr* hdr, int npeers, size_t numel, int scale,
                               float* out, oracle_hdr* out_hdr) {
  oracle_hdr tot = {0, 0, 0, 0};
  for (int r = 0; r < npeers; ++r) {
    tot.num_gradients += hdr[r].num_gradients;
    tot.num_skipped += hdr[r].num_skipped;
    tot.batch_size += hdr[r].batch_size;
    tot.has_grads += in[r] != NULL;
  }
  const int do_scale = scale && tot.num_gradients;
  const float s = do_scale ? 1.0f / (float)tot.num_gradients : 1.0f;
  for (size_t i = 0; i < numel; ++i) {
    float acc = 0.0f;
    int first = 1;
    for (int r = 0; r < npeers; ++r) {
      if (!in[r]) continue;
      if (first) {
        acc = in[r][i];
        first = 0;
      } else {
        acc = acc + in[r][i];
      }
    }
    out[i] = do_scale ? acc * s : acc;
  }
  if (out_hdr) *out_hdr = tot;
  return 0;
}

/* ---- learner-side steps next to the hot paths ------------------------------------------ */

static float oracle_clamp_max(float x, float c, int on) {
  if (!on || x != x) return x; /* torch.clamp propagates NaN */
  return x < c ? x : c;
}

/* V-trace from log importance weights: examples/common/vtrace.py:207-242 from_importance_weights, [T, B] fp32.
 *   rhos = exp(log_rhos); clipped_rhos = clamp(rhos, max=clip_rho); cs = clamp(rhos, max=1)            (:207-213)
 *   deltas = clipped_rhos * (rewards + discounts * values_t_plus_1 - values)                              (:215-219)
 *   acc = deltas[t] + discounts[t] * cs[t] * acc   for t = T-1 .. 0;   vs = acc + values                  (:221-230)
 *   pg_advantages = clamp(rhos, max=clip_pg_rho) * (rewards + discounts * vs_t_plus_1 - values)            (:233-239)
 * Every operation rounds to fp32 on its own (compile with -ffp-contract=off).  expf is the C library's: it may differ
 * from the device's expf in the last bit, which is why parity with this function is checked at 1e-6 relative while
 * parity with the PyTorch restatement on the same device is bit-exact. */
int oracle_vtrace(const float* log_rhos, const float* discounts, const float* rewards, const float* values,
                  const float* bootstrap, int has_clip_rho, float clip_rho, int has_clip_pg_rho, float clip_pg_rho,
                  size_t T, size_t B, float* vs_out, float* pg_out) {
  for (size_t j = 0; j < B; ++j) {
    float acc = 0.0f, v_next = bootstrap[j], vs_next = bootstrap[j];
    for (size_t t = T; t-- > 0;) {
      const size_t i = t * B + j;
      const float rho = expf(log_rhos[i]);
      const float d = discounts[i], r = rewards[i], v = values[i];
      const float crho = oracle_clamp_max(rho, clip_rho, has_clip_rho);
      const float c = oracle_clamp_max(rho, 1.0f, 1);
      float tmp = d * v_next;
      tmp = r + tmp;
      tmp = tmp - v;
      const float delta = crho * tmp;
      float dc = d * c;
      dc = dc * acc;
      acc = delta + dc;
      const float vs = acc + v;
      float q = d * vs_next;
      q = r + q;
      q = q - v;
      pg_out[i] = oracle_clamp_max(rho, clip_pg_rho, has_clip_pg_rho) * q;
      vs_out[i] = vs;
      v_next = v;
      vs_next = vs;
    }
  }
  return 0;
}

/* x.float() / 255.0 as ATen evaluates it on CUDA: fp32 multiplication by the fp32 reciprocal of the scalar
 * (examples/atari/models.py:94; aten/src/ATen/native/cuda/BinaryDivTrueKernel.cu: `inv_b = 1 / b`, `a * inv_b`). */
int oracle_u8_to_f32(const uint8_t* src, float* dst, size_t n, float scale) {
  for (size_t i = 0; i < n; ++i) dst[i] = (float)src[i] * scale;
  return 0;
}
