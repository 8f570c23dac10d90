// Parameter-arena and activation-stash layouts of one SET network ("net" = one reference
// TransformerModel, src/SEActor.py:170-287).  Host-only, plain C++.
//
// Every tensor of the reference state_dict has a slot in one flat
// fp32 arena; the Python nn.Parameters are views into it (so checkpoints load and the
// optimizer/all-reduce stream one buffer).  Tensors that the kernels consume as one
// stacked matrix (q|k|v, linear3|linear1) are placed adjacently.  Live tensors first,
// the dead nn.MultiheadAttention leftovers (in_proj_*, out_proj.*) last, so Adam and the
// gradient all-reduce touch only the live prefix.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace sgrl {

constexpr int D = 128;        // embedding width
constexpr int HEADS = 2;      // hard-wired in the reference (subequivariant_attentions.py:117)
constexpr int HD = 128;       // per-head width of q/k/v (2*D/HEADS)
constexpr int HID = 256;      // feed-forward width
constexpr int CH = 32;        // invariant channels (30 learned + gravity + direction)
constexpr int NPJ = 30;       // learned projection channels
constexpr int GN = 8;         // 3-vectors per limb observation
constexpr int MAXN = 16;      // max limbs per graph the attention kernel supports (tables: 15)
constexpr int MAX_NODE = 15;  // rows of each positional table
constexpr int MAX_LAYERS = 8;
constexpr int ALIGN = 16;     // floats (64 B): keeps every tensor 16B-aligned for float4/TMA
// The Gram G = Z^T Z is symmetric, so vec(G) (1024 floats per token) is stored and contracted as its upper triangle:
// GP = 528 entries zero-padded to GP_K = 544 = 17 k-blocks of 32.  The three weights that consume vec(G)
// (self_attn.linear_g1, linear_g1, linear1_g) are folded once per pass into
// W'[o][p(i,j)] = W[o][32i+j] + W[o][32j+i] (i<j), W[o][33i] (i=j): the same contraction with 47 % fewer MACs and
// 1.9 KB instead of 4 KB of HBM traffic per token and Gram.  Parameters, gradients and state_dict stay (rows,1024).
//
// Packing order p(i,j), i <= j — chosen so that a k-block of 32 entries is cheap to GENERATE inside the consuming GEMM
// (gemm_tc.cuh, GRAM operand): the 32 channels are cut into 8 blocks of 4;
//   k-blocks 0..13  : the 28 off-diagonal block pairs (Ib < Jb) in lexicographic order, two per k-block, each 4x4
//                     row-major (16 entries from 2 x 12 values of Z);
//   k-blocks 14..16 : the 8 diagonal blocks, three per k-block, each its 10 entries (a <= b) row-major; the slots
//                     30, 31 of k-blocks 14, 15 and 20..31 of k-block 16 are zero padding.
constexpr int GP = CH * (CH + 1) / 2;   // 528 real entries
constexpr int GP_K = 544;
constexpr int GP_OFF = 448;             // first entry of the diagonal blocks (28 pairs x 16)
__host__ __device__ constexpr int tri_index(int i, int j) {   // i <= j
  const int ib = i >> 2, jb = j >> 2, a = i & 3, b = j & 3;
  if (ib < jb) return (ib * 7 - (ib * (ib - 1)) / 2 + (jb - ib - 1)) * 16 + a * 4 + b;
  return GP_OFF + (ib / 3) * 32 + (ib % 3) * 10 + (a * 4 - (a * (a - 1)) / 2 + (b - a));
}

enum Kind { ACTOR = 0, CRITIC = 1 };

// ---- per-layer parameter ids -------------------------------------------------------
enum LP {
  L_GPROJ, L_G1_W, L_G1_B, L_G2_W, L_G2_B,
  L_Q_W, L_K_W, L_V_W, L_Q_B, L_K_B, L_V_B,
  L_VG_W, L_NGO_W, L_NGO_B, L_GO_W, L_N1_W, L_N1_B,
  L_GP2, L_GP3, L_FG1_W, L_FG1_B, L_FG2_W, L_FG2_B,
  L_L3_W, L_L1_W, L_L3_B, L_L1_B, L_L4_W, L_L4_B, L_L5_W, L_L2_W, L_L2_B, L_N2_W, L_N2_B,
  LP_COUNT
};
// ---- global parameter ids ----------------------------------------------------------
enum GP {
  G_POS0, G_POS1, G_POS2, G_NORM_W, G_NORM_B, G_REL_W, G_REL_B, G_GENC_W, G_ENC_W, G_ENC_B,
  G_GG_W, G_GPH_W /*actor only: head g_proj*/, G_H1G_W, G_H1G_B, G_H2G_W, G_H2G_B,
  G_H1NG_W, G_H1NG_B, G_H2NG_W, G_H2NG_B,
  G_DNG_W, G_DNG_B,                       // critic only
  G_DG_W, G_H1M_W, G_H1M_B, G_H2M_W, G_H2M_B,  // actor only
  GP_COUNT
};
// dead per-layer ids
enum DP { D_INW, D_INB, D_OUTW, D_OUTB, DP_COUNT };

struct ParamRow { const char* name; int rows; int cols; };

inline const ParamRow* layer_rows() {
  static const ParamRow r[LP_COUNT] = {
      {"self_attn.g_proj.weight", NPJ, D},
      {"self_attn.linear_g1.weight", 2 * D, CH * CH}, {"self_attn.linear_g1.bias", 2 * D, 0},
      {"self_attn.linear_g2.weight", D, 2 * D}, {"self_attn.linear_g2.bias", D, 0},
      {"self_attn.q_proj.weight", 2 * D, 2 * D}, {"self_attn.k_proj.weight", 2 * D, 2 * D},
      {"self_attn.v_proj.weight", 2 * D, 2 * D},
      {"self_attn.q_proj.bias", 2 * D, 0}, {"self_attn.k_proj.bias", 2 * D, 0}, {"self_attn.v_proj.bias", 2 * D, 0},
      {"self_attn.vg_proj.weight", 2 * D - 2 * HEADS, D},
      {"self_attn.ng_out.weight", D, 2 * D}, {"self_attn.ng_out.bias", D, 0},
      {"self_attn.g_out.weight", D, 2 * D},
      {"norm1.weight", D, 0}, {"norm1.bias", D, 0},
      {"g_proj2.weight", NPJ, D}, {"g_proj3.weight", NPJ, D},
      {"linear_g1.weight", HID, CH * CH}, {"linear_g1.bias", HID, 0},
      {"linear_g2.weight", D, HID}, {"linear_g2.bias", D, 0},
      {"linear3.weight", HID, 2 * D}, {"linear1.weight", HID, 2 * D},
      {"linear3.bias", HID, 0}, {"linear1.bias", HID, 0},
      {"linear4.weight", CH * CH, HID}, {"linear4.bias", CH * CH, 0},
      {"linear5.weight", D, CH},
      {"linear2.weight", D, HID}, {"linear2.bias", D, 0},
      {"norm2.weight", D, 0}, {"norm2.bias", D, 0},
  };
  return r;
}

inline const ParamRow* dead_rows() {
  static const ParamRow r[DP_COUNT] = {
      {"self_attn.in_proj_weight", 3 * D, D}, {"self_attn.in_proj_bias", 3 * D, 0},
      {"self_attn.out_proj.weight", D, D}, {"self_attn.out_proj.bias", D, 0},
  };
  return r;
}

// rows/cols of global params depend on kind; cols==0 means 1-D; rows==0 means "absent for this kind"
inline ParamRow global_row(int kind, int id) {
  const int ng = (kind == ACTOR ? 41 : 44) - 3 * GN;  // invariant scalars per limb: 17 | 20
  switch (id) {
    case G_POS0: return {"pos_encoder.embeddings.0.weight", MAX_NODE, D / 3};
    case G_POS1: return {"pos_encoder.embeddings.1.weight", MAX_NODE, D / 3};
    case G_POS2: return {"pos_encoder.embeddings.2.weight", MAX_NODE, D / 3 + D % 3};
    case G_NORM_W: return {"transformer_encoder.norm.weight", D, 0};
    case G_NORM_B: return {"transformer_encoder.norm.bias", D, 0};
    case G_REL_W: return {"transformer_encoder.rel_encoder.weight", HEADS, 3};
    case G_REL_B: return {"transformer_encoder.rel_encoder.bias", HEADS, 0};
    case G_GENC_W: return {"g_encoder.weight", D, GN};
    case G_ENC_W: return {"encoder.weight", D, ng};
    case G_ENC_B: return {"encoder.bias", D, 0};
    case G_GG_W: return {"gg_proj.weight", NPJ, D + GN};
    case G_GPH_W: return {"g_proj.weight", kind == ACTOR ? NPJ : 0, D + GN};
    case G_H1G_W: return {"linear1_g.weight", D, CH * CH};
    case G_H1G_B: return {"linear1_g.bias", D, 0};
    case G_H2G_W: return {"linear2_g.weight", D, D};
    case G_H2G_B: return {"linear2_g.bias", D, 0};
    case G_H1NG_W: return {"linear1_ng.weight", D, D + ng};
    case G_H1NG_B: return {"linear1_ng.bias", D, 0};
    case G_H2NG_W: return {"linear2_ng.weight", D, D};
    case G_H2NG_B: return {"linear2_ng.bias", D, 0};
    case G_DNG_W: return {"decoder_ng.weight", kind == CRITIC ? 1 : 0, 2 * D};
    case G_DNG_B: return {"decoder_ng.bias", kind == CRITIC ? 1 : 0, 0};
    case G_DG_W: return {"decoder_g.weight", kind == ACTOR ? 1 : 0, CH};
    case G_H1M_W: return {"linear1_m.weight", kind == ACTOR ? 2 * D : 0, 2 * D};
    case G_H1M_B: return {"linear1_m.bias", kind == ACTOR ? 2 * D : 0, 0};
    case G_H2M_W: return {"linear2_m.weight", kind == ACTOR ? CH * CH : 0, 2 * D};
    case G_H2M_B: return {"linear2_m.bias", kind == ACTOR ? CH * CH : 0, 0};
  }
  return {"", 0, 0};
}

inline long long numel(const ParamRow& r) { return (long long)r.rows * (r.cols ? r.cols : 1); }
inline long long align_up(long long x, long long a = ALIGN) { return (x + a - 1) / a * a; }

struct NetLayout {
  int kind, n_layers;
  long long lp[MAX_LAYERS][LP_COUNT];
  long long gp[GP_COUNT];       // -1 when absent
  long long dp[MAX_LAYERS][DP_COUNT];   // relative to the start of the dead region
  long long live_floats;        // floats of live tensors (= z-stride between twin critics)
  long long dead_floats;        // floats of dead tensors of one net
};
// Arena of an nb-net module: [live net0 | live net1 | ... | dead net0 | dead net1 | ...].
// Kernels only ever see the live prefix (gradients, Adam, all-reduce run over
// nb*live_floats); the dead tail exists for state_dict compatibility and Polyak.

inline NetLayout make_layout(int kind, int n_layers) {
  NetLayout L;
  L.kind = kind; L.n_layers = n_layers;
  long long off = 0;
  for (int l = 0; l < n_layers; ++l)
    for (int i = 0; i < LP_COUNT; ++i) { L.lp[l][i] = off; off = align_up(off + numel(layer_rows()[i])); }
  for (int i = 0; i < GP_COUNT; ++i) {
    ParamRow r = global_row(kind, i);
    if (r.rows == 0) { L.gp[i] = -1; continue; }
    L.gp[i] = off; off = align_up(off + numel(r));
  }
  L.live_floats = align_up(off, 64);
  off = 0;
  for (int l = 0; l < n_layers; ++l)
    for (int i = 0; i < DP_COUNT; ++i) { L.dp[l][i] = off; off = align_up(off + numel(dead_rows()[i])); }
  L.dead_floats = align_up(off, 64);
  return L;
}

// Enumerate every tensor with its reference state_dict name (relative to the TransformerModel).
struct ParamInfo { char name[160]; int rows, cols; long long offset; int live; };

inline int enumerate_params(int kind, int n_layers, ParamInfo* out, int cap) {
  NetLayout L = make_layout(kind, n_layers);
  int n = 0;
  auto put = [&](const char* nm, int rows, int cols, long long off, int live) {
    if (out && n < cap) {
      snprintf(out[n].name, sizeof(out[n].name), "%s", nm);
      out[n].rows = rows; out[n].cols = cols; out[n].offset = off; out[n].live = live;
    }
    ++n;
  };
  char buf[160];
  for (int l = 0; l < n_layers; ++l) {
    for (int i = 0; i < LP_COUNT; ++i) {
      snprintf(buf, sizeof(buf), "transformer_encoder.layers.%d.%s", l, layer_rows()[i].name);
      put(buf, layer_rows()[i].rows, layer_rows()[i].cols, L.lp[l][i], 1);
    }
  }
  for (int i = 0; i < GP_COUNT; ++i) {
    ParamRow r = global_row(kind, i);
    if (r.rows == 0) continue;
    put(r.name, r.rows, r.cols, L.gp[i], 1);
  }
  for (int l = 0; l < n_layers; ++l) {
    for (int i = 0; i < DP_COUNT; ++i) {
      snprintf(buf, sizeof(buf), "transformer_encoder.layers.%d.%s", l, dead_rows()[i].name);
      put(buf, dead_rows()[i].rows, dead_rows()[i].cols, L.dp[l][i], 0);
    }
  }
  return n;
}

// ---- activation stash (forward -> backward) -----------------------------------------
// Sizes in floats per token.  Per-layer buffers are replicated n_layers times when the
// forward runs with keep=1 (training) and aliased to one layer when keep=0 (rollout).
enum LS {
  S_Z1, S_F1, S_G1, S_A1, S_UA, S_QKV, S_VGP, S_P, S_O, S_OG, S_X1, S_ST1, S_DV,
  S_Z2, S_Z3, S_G2, S_F2, S_A2, S_UB, S_T31, S_MM, S_R, S_FF, S_X2, S_ST2, S_VGIN,
  LS_COUNT
};
inline const int* layer_stash_sizes() {
  static const int s[LS_COUNT] = {
      96, 1, GP_K, 256, 256, 768, 756, HEADS * MAXN, 256, 768, 128, 2, 384,
      96, 96, GP_K, 1, 256, 256, 512, 1024, 96, 128, 128, 2, 384};
  return s;
}
inline const char* const* layer_stash_names() {
  static const char* const n[LS_COUNT] = {
      "Z1", "F1", "G1", "A1", "UA", "QKV", "VGP", "P", "O", "OG", "X1", "ST1", "DV",
      "Z2", "Z3", "G2", "F2", "A2", "UB", "T31", "MM", "R", "FF", "X2", "ST2", "VGIN"};
  return n;
}
enum GS {
  T_V0, T_GD, T_SH, T_HL, T_STF, T_VGF, T_ZH, T_ZH2, T_GH, T_FH, T_AH, T_UH, T_BH, T_M1, T_MH, T_RH, T_W3, T_OUT,
  GS_COUNT
};
inline int global_stash_size(int kind, int id) {
  const int ks = D + ((kind == ACTOR ? 41 : 44) - 3 * GN);
  switch (id) {
    case T_V0: return 24; case T_GD: return 6; case T_SH: return ks; case T_HL: return 128;
    case T_STF: return 2; case T_VGF: return 384; case T_ZH: return 96; case T_ZH2: return kind == ACTOR ? 96 : 0;
    case T_GH: return GP_K; case T_FH: return 1; case T_AH: return 128; case T_UH: return 256; case T_BH: return 128;
    case T_M1: return kind == ACTOR ? 256 : 0; case T_MH: return kind == ACTOR ? 1024 : 0;
    case T_RH: return kind == ACTOR ? 96 : 0; case T_W3: return kind == ACTOR ? 3 : 0;
    case T_OUT: return kind == ACTOR ? 3 : 1;
  }
  return 0;
}
inline const char* const* global_stash_names() {
  static const char* const n[GS_COUNT] = {"V0", "GD", "SH", "HL", "STF", "VGF", "ZH", "ZH2", "GH", "FH", "AH", "UH", "BH", "M1", "MH", "RH", "W3", "OUT"};
  return n;
}

// Folded (triangle-packed) copies of the vec(G) consumers, rebuilt at the start of every forward into the stash:
// per layer [self_attn.linear_g1 (256 x GP_K) | linear_g1 (256 x GP_K)], then linear1_g (128 x GP_K); three planes
// (fp32, tf32 hi, tf32 lo) of fold_floats() each.
inline long long fold_offset(int n_layers, int l, int which) {       // l == n_layers: the head's linear1_g
  return l < n_layers ? (long long)(2 * l + which) * HID * GP_K : (long long)2 * n_layers * HID * GP_K;
}
inline long long fold_floats(int n_layers) { return align_up((long long)2 * n_layers * HID * GP_K + (long long)D * GP_K, 64); }

struct StashLayout {
  long long ls[MAX_LAYERS][LS_COUNT];
  long long gs[GS_COUNT];
  long long wf;       // folded weights: 3 planes of wf_plane floats (independent of T)
  long long wf_plane;
  long long total;    // floats per net instance (z-stride)
};

inline StashLayout make_stash(int kind, int n_layers, long long T, int keep) {
  StashLayout S;
  long long off = 0;
  auto take = [&](long long per_tok) { long long o = off; off = align_up(off + per_tok * T, 64); return o; };
  for (int l = 0; l < n_layers; ++l) {
    if (l > 0 && !keep) { for (int i = 0; i < LS_COUNT; ++i) S.ls[l][i] = S.ls[0][i]; continue; }
    for (int i = 0; i < LS_COUNT; ++i) S.ls[l][i] = take(layer_stash_sizes()[i]);
  }
  if (!keep && n_layers > 1) {
    // layer l reads UA/VGIN of layer l while layer l writes UA/VGIN of l+1: ping-pong those two
    long long ua2 = take(layer_stash_sizes()[S_UA]), vg2 = take(layer_stash_sizes()[S_VGIN]);
    for (int l = 1; l < n_layers; l += 2) { S.ls[l][S_UA] = ua2; S.ls[l][S_VGIN] = vg2; }
  }
  for (int i = 0; i < GS_COUNT; ++i) S.gs[i] = take(global_stash_size(kind, i));
  S.wf_plane = fold_floats(n_layers);
  S.wf = off; off += 3 * S.wf_plane;
  S.total = off;
  return S;
}

}  // namespace sgrl
