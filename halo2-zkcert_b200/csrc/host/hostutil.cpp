#include "hostutil.h"

namespace zkc {
namespace host {

// ---- Blake2b ------------------------------------------------------------------------------------------------
static const uint64_t B2_IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                  0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
static const uint8_t B2_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
static inline uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
static inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

void Blake2b::init(const char personal[16]) {
  for (int i = 0; i < 8; ++i) h[i] = B2_IV[i];
  // parameter block: digest_length = 64, key_length = 0, fanout = 1, depth = 1; personal at bytes 48..63
  h[0] ^= 0x01010000ULL ^ 64ULL;
  h[6] ^= load64((const uint8_t*)personal);
  h[7] ^= load64((const uint8_t*)personal + 8);
  t[0] = t[1] = 0;
  buflen = 0;
  memset(buf, 0, sizeof buf);
}
void Blake2b::compress(const uint8_t block[128], bool last) {
  uint64_t m[16], v[16];
  for (int i = 0; i < 16; ++i) m[i] = load64(block + 8 * i);
  for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = B2_IV[i]; }
  v[12] ^= t[0]; v[13] ^= t[1];
  if (last) v[14] = ~v[14];
#define B2G(r, i, a, b, c, d)                         \
  a = a + b + m[B2_SIGMA[r][2 * i]]; d = rotr64(d ^ a, 32); c = c + d; b = rotr64(b ^ c, 24); \
  a = a + b + m[B2_SIGMA[r][2 * i + 1]]; d = rotr64(d ^ a, 16); c = c + d; b = rotr64(b ^ c, 63);
  for (int r = 0; r < 12; ++r) {
    B2G(r, 0, v[0], v[4], v[8], v[12]); B2G(r, 1, v[1], v[5], v[9], v[13]);
    B2G(r, 2, v[2], v[6], v[10], v[14]); B2G(r, 3, v[3], v[7], v[11], v[15]);
    B2G(r, 4, v[0], v[5], v[10], v[15]); B2G(r, 5, v[1], v[6], v[11], v[12]);
    B2G(r, 6, v[2], v[7], v[8], v[13]); B2G(r, 7, v[3], v[4], v[9], v[14]);
  }
#undef B2G
  for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
}
void Blake2b::update(const uint8_t* in, size_t len) {
  while (len > 0) {
    if (buflen == 128) {   // buffer full and more input follows: it is not the last block
      t[0] += 128; if (t[0] < 128) t[1]++;
      compress(buf, false);
      buflen = 0;
    }
    size_t take = 128 - buflen;
    if (take > len) take = len;
    memcpy(buf + buflen, in, take);
    buflen += take; in += take; len -= take;
  }
}
void Blake2b::finalize(uint8_t out[64]) const {
  Blake2b c = *this;
  c.t[0] += c.buflen; if (c.t[0] < c.buflen) c.t[1]++;
  memset(c.buf + c.buflen, 0, 128 - c.buflen);
  c.compress(c.buf, true);
  memcpy(out, c.h, 64);
}

// ---- Keccak-256 ---------------------------------------------------------------------------------------------
static inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static void keccak_f(uint64_t st[25]) {
  static const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
                                  0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
                                  0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
                                  0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
                                  0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
                                  0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};  // [x + 5y]
  for (int round = 0; round < 24; ++round) {
    uint64_t C[5], D[5], B[25];
    for (int x = 0; x < 5; ++x) C[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
    for (int x = 0; x < 5; ++x) D[x] = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
    for (int i = 0; i < 25; ++i) st[i] ^= D[i % 5];
    for (int x = 0; x < 5; ++x)
      for (int y = 0; y < 5; ++y) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(st[x + 5 * y], ROT[x + 5 * y]);
    for (int x = 0; x < 5; ++x)
      for (int y = 0; y < 5; ++y) st[x + 5 * y] = B[x + 5 * y] ^ ((~B[(x + 1) % 5 + 5 * y]) & B[(x + 2) % 5 + 5 * y]);
    st[0] ^= RC[round];
  }
}
void keccak256(const uint8_t* in, size_t len, uint8_t out[32]) {
  const size_t rate = 136;
  uint64_t st[25];
  memset(st, 0, sizeof st);
  while (len >= rate) {
    for (size_t i = 0; i < rate / 8; ++i) st[i] ^= load64(in + 8 * i);
    keccak_f(st);
    in += rate; len -= rate;
  }
  uint8_t last[136];
  memset(last, 0, sizeof last);
  memcpy(last, in, len);
  last[len] ^= 0x01;
  last[rate - 1] ^= 0x80;
  for (size_t i = 0; i < rate / 8; ++i) st[i] ^= load64(last + 8 * i);
  keccak_f(st);
  memcpy(out, st, 32);
}

// ---- ChaCha20 -----------------------------------------------------------------------------------------------
static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
ChaCha20Rng::ChaCha20Rng(const uint8_t seed[32], int dr) : counter(0), pos(16), double_rounds(dr) { memcpy(key, seed, 32); }
static void chacha_block(const uint32_t key[8], uint64_t counter, uint32_t out[16], int double_rounds) {
  uint32_t s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                    (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
  uint32_t x[16];
  memcpy(x, s, sizeof x);
#define QR(a, b, c, d) \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
  for (int i = 0; i < double_rounds; ++i) {
    QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
    QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
  }
#undef QR
  for (int i = 0; i < 16; ++i) out[i] = x[i] + s[i];
}
uint32_t ChaCha20Rng::next_u32() {
  if (pos >= 16) { chacha_block(key, counter++, block, double_rounds); pos = 0; }
  return block[pos++];
}
uint64_t ChaCha20Rng::next_u64() {
  uint64_t lo = next_u32();
  uint64_t hi = next_u32();
  return lo | (hi << 32);
}
Fr ChaCha20Rng::fr_random() {
  uint8_t b[64];
  for (int i = 0; i < 8; ++i) { uint64_t v = next_u64(); memcpy(b + 8 * i, &v, 8); }
  return fr_from_u512_le(b);
}
void ChaCha20Rng::seek(uint64_t word) {
  counter = word / 16; pos = 16;
  if (word % 16) { chacha_block(key, counter++, block, double_rounds); pos = (int)(word % 16); }
}
void ChaCha20Rng::fill_bytes(uint8_t* out, size_t nbytes) {
  for (size_t i = 0; i + 4 <= nbytes; i += 4) { const uint32_t w = next_u32(); memcpy(out + i, &w, 4); }
}
void seed_from_u64(uint64_t state, uint8_t seed[32]) {
  const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
  for (int i = 0; i < 8; ++i) {
    state = state * MUL + INC;
    uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    memcpy(seed + 4 * i, &x, 4);
  }
}

// ---- scalars ------------------------------------------------------------------------------------------------
static Fr reduce_256(const uint8_t b[32]) {   // raw 256-bit integer -> canonical representative < r (as a plain limb vector)
  Fr v; memcpy(v.v, b, 32);
  while (geq_mod<FrP>(v.v)) sub_mod_inplace<FrP>(v.v);
  return v;
}
Fr fr_from_u512_le(const uint8_t bytes[64]) {
  // value = lo + hi * 2^256;  Montgomery(lo) = lo * R^2 / R,  Montgomery(hi * 2^256) = hi * R^3 / R
  const Fr lo = reduce_256(bytes), hi = reduce_256(bytes + 32);
  const Fr r2 = fe_r2<FrP>();
  const Fr r3 = fe_mul(r2, r2);
  return fe_add(fe_mul(lo, r2), fe_mul(hi, r3));
}
void fr_to_repr(const Fr& a, uint8_t out[32]) { Fr c = fe_to_canonical(a); memcpy(out, c.v, 32); }
void fq_to_repr(const Fq& a, uint8_t out[32]) { Fq c = fe_to_canonical(a); memcpy(out, c.v, 32); }
int fr_cmp_canonical(const Fr& a, const Fr& b) {
  Fr x = fe_to_canonical(a), y = fe_to_canonical(b);
  for (int i = 7; i >= 0; --i) { if (x.v[i] < y.v[i]) return -1; if (x.v[i] > y.v[i]) return 1; }
  return 0;
}

// ---- Poseidon ---------------------------------------------------------------------------------------------------
namespace {
struct Grain {
  bool s[80];
  int head = 0;   // ring buffer start
  Grain() {
    int n = 0;
    auto append = [&](int bits, uint64_t v) { for (int i = bits - 1; i >= 0; --i) s[n++] = (v >> i) & 1; };
    append(2, 1); append(4, 0); append(12, 254); append(12, 3); append(10, 8); append(10, 57); append(30, (1ull << 30) - 1);
    for (int i = 0; i < 160; ++i) new_bit();
  }
  bool at(int i) const { return s[(head + i) % 80]; }
  bool new_bit() {
    const bool b = at(62) ^ at(51) ^ at(38) ^ at(23) ^ at(13) ^ at(0);
    s[head] = b;              // drop the oldest bit, append the new one
    head = (head + 1) % 80;
    return b;
  }
  bool next_bit() {
    bool b = new_bit();
    while (!b) { new_bit(); b = new_bit(); }
    return new_bit();
  }
  void take(uint8_t repr[32]) {   // 254 bits, MSB first, into a little-endian byte string
    memset(repr, 0, 32);
    for (int i = 0; i < 254; ++i) { const int p = 253 - i; if (next_bit()) repr[p / 8] |= (uint8_t)(1u << (p % 8)); }
  }
  Fr next_field_element() {
    for (;;) {
      uint8_t repr[32]; take(repr);
      Fr v; memcpy(v.v, repr, 32);
      if (!geq_mod<FrP>(v.v)) return fe_from_canonical(v);
    }
  }
  Fr next_field_element_without_rejection() {
    uint8_t repr[32]; take(repr);
    Fr v; memcpy(v.v, repr, 32);
    while (geq_mod<FrP>(v.v)) sub_mod_inplace<FrP>(v.v);
    return fe_from_canonical(v);
  }
};
struct PoseidonSpec {
  Fr constants[65][3];
  Fr mds[3][3];
  PoseidonSpec() {
    Grain g;
    for (int r = 0; r < 65; ++r) for (int i = 0; i < 3; ++i) constants[r][i] = g.next_field_element();
    Fr xs[3], ys[3];
    for (int i = 0; i < 3; ++i) xs[i] = g.next_field_element_without_rejection();
    for (int i = 0; i < 3; ++i) ys[i] = g.next_field_element_without_rejection();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) mds[i][j] = fe_inv(fe_add(xs[i], ys[j]));
  }
};
const PoseidonSpec& the_spec() { static const PoseidonSpec sp; return sp; }
Fr pow5(const Fr& a) { Fr a2 = fe_sqr(a); return fe_mul(fe_sqr(a2), a); }
void poseidon_permute(Fr st[3]) {
  const PoseidonSpec& sp = the_spec();
  for (int r = 0; r < 65; ++r) {
    for (int i = 0; i < 3; ++i) st[i] = fe_add(st[i], sp.constants[r][i]);
    if (r < 4 || r >= 61) { for (int i = 0; i < 3; ++i) st[i] = pow5(st[i]); }
    else st[0] = pow5(st[0]);
    Fr o[3];
    for (int i = 0; i < 3; ++i) o[i] = fe_add(fe_add(fe_mul(sp.mds[i][0], st[0]), fe_mul(sp.mds[i][1], st[1])), fe_mul(sp.mds[i][2], st[2]));
    st[0] = o[0]; st[1] = o[1]; st[2] = o[2];
  }
}
}  // namespace

void poseidon_spec(const Fr** constants, const Fr** mds) { *constants = &the_spec().constants[0][0]; *mds = &the_spec().mds[0][0]; }

PoseidonSponge::PoseidonSponge() {
  Fr c = fe_zero<FrP>(); c.v[2] = 1;    // 2^64
  state[0] = fe_from_canonical(c); state[1] = fe_zero<FrP>(); state[2] = fe_zero<FrP>();
}
void PoseidonSponge::absorb(const Fr* chunk, size_t len) {
  for (size_t i = 0; i < len; ++i) state[1 + i] = fe_add(state[1 + i], chunk[i]);
  if (len < 2) state[1 + len] = fe_add(state[1 + len], fe_one<FrP>());
  poseidon_permute(state);
}
Fr PoseidonSponge::squeeze() {
  std::vector<Fr> b; b.swap(buf);
  const bool exact = b.size() % 2 == 0;
  for (size_t i = 0; i < b.size(); i += 2) absorb(b.data() + i, std::min<size_t>(2, b.size() - i));
  if (exact) absorb(nullptr, 0);
  return state[1];
}

// ---- transcripts --------------------------------------------------------------------------------------------
Transcript::Transcript(int kind_, int pf) : kind(kind_), point_format(pf) {
  if (kind == 0) b2.init("Halo2-Transcript");
}
void Transcript::absorb(const uint8_t* p, size_t n) {
  if (kind == 0) b2.update(p, n);
  else kbuf.insert(kbuf.end(), p, p + n);
}
static void reverse32(uint8_t b[32]) { for (int i = 0; i < 16; ++i) { uint8_t t = b[i]; b[i] = b[31 - i]; b[31 - i] = t; } }

Fr Transcript::squeeze_challenge() {
  if (kind == 3) return pos.squeeze();
  if (kind == 2) {
    // EvmTranscript: keccak256(buf ++ [1 if len == 32]); the digest replaces the buffer; challenge = digest (BE) mod r
    std::vector<uint8_t> t = kbuf;
    if (t.size() == 32) t.push_back(1);
    uint8_t h[32];
    keccak256(t.data(), t.size(), h);
    kbuf.assign(h, h + 32);
    uint8_t wide[64];
    memset(wide, 0, sizeof wide);
    for (int i = 0; i < 32; ++i) wide[i] = h[31 - i];
    return fr_from_u512_le(wide);
  }
  const uint8_t prefix = 0;
  absorb(&prefix, 1);
  uint8_t d[64];
  if (kind == 0) {
    b2.finalize(d);
  } else {
    std::vector<uint8_t> t = kbuf;
    t.push_back(10);
    keccak256(t.data(), t.size(), d);
    t.back() = 11;
    keccak256(t.data(), t.size(), d + 32);
  }
  return fr_from_u512_le(d);
}
int Transcript::common_point(const G1Affine& p) {
  if (affine_is_identity(p)) return 1;
  if (kind == 3) {   // coordinates reduced from Fq into Fr (fe_to_fe)
    uint8_t xb[32], yb[32];
    fq_to_repr(p.x, xb); fq_to_repr(p.y, yb);
    Fr e[2] = {fe_from_canonical(reduce_256(xb)), fe_from_canonical(reduce_256(yb))};
    pos.update(e, 2);
    return 0;
  }
  if (kind == 2) {
    uint8_t b[64];
    fq_to_repr(p.x, b); fq_to_repr(p.y, b + 32);
    reverse32(b); reverse32(b + 32);
    absorb(b, 64);
    return 0;
  }
  uint8_t b[65];
  b[0] = 1;
  fq_to_repr(p.x, b + 1);
  fq_to_repr(p.y, b + 33);
  absorb(b, 65);
  return 0;
}
void Transcript::common_scalar(const Fr& s) {
  if (kind == 3) { pos.update(&s, 1); return; }
  if (kind == 2) {
    uint8_t b[32];
    fr_to_repr(s, b);
    reverse32(b);
    absorb(b, 32);
    return;
  }
  uint8_t b[33];
  b[0] = 2;
  fr_to_repr(s, b + 1);
  absorb(b, 33);
}
int Transcript::write_point(const G1Affine& p) {
  if (common_point(p)) return 1;
  uint8_t xb[32], yb[32];
  fq_to_repr(p.x, xb);
  fq_to_repr(p.y, yb);
  if (kind == 2) {
    reverse32(xb); reverse32(yb);
    proof.insert(proof.end(), xb, xb + 32);
    proof.insert(proof.end(), yb, yb + 32);
    return 0;
  }
  const uint8_t sign = yb[0] & 1;
  xb[31] |= point_format == 0 ? (uint8_t)(sign << 7) : (uint8_t)(sign << 6);
  proof.insert(proof.end(), xb, xb + 32);
  return 0;
}
void Transcript::write_scalar(const Fr& s) {
  common_scalar(s);
  uint8_t b[32];
  fr_to_repr(s, b);
  if (kind == 2) reverse32(b);
  proof.insert(proof.end(), b, b + 32);
}


// ---- reader side ---------------------------------------------------------------------------------------------------
bool g1_decompress(const Fq& x, int y_is_odd, Fq* y) {
  static const unsigned long long SQRT_EXP[4] = {0x4f082305b61f3f52ull, 0x65e05aa45a1c72a3ull, 0x6e14116da0605617ull, 0x0c19139cb84c680aull};   // (p + 1) / 4
  Fq three = fe_zero<FqP>(); three.v[0] = 3; three = fe_from_canonical(three);
  const Fq y2 = fe_add(fe_mul(fe_sqr(x), x), three);
  Fq acc = fe_one<FqP>();
  for (int w = 3; w >= 0; --w)
    for (int b = 63; b >= 0; --b) { acc = fe_sqr(acc); if ((SQRT_EXP[w] >> b) & 1) acc = fe_mul(acc, y2); }
  if (!fe_eq(fe_sqr(acc), y2)) return false;
  if ((int)(fe_to_canonical(acc).v[0] & 1) != y_is_odd) acc = fe_neg(acc);
  *y = acc;
  return true;
}
static bool fq_from_repr(const uint8_t b[32], Fq* out) {   // canonical little-endian, must be < p
  Fq v; memcpy(v.v, b, 32);
  if (geq_mod<FqP>(v.v)) return false;
  *out = fe_from_canonical(v);
  return true;
}
int Transcript::read_point(G1Affine* out) {
  G1Affine p;
  if (kind == 2) {   // 64 bytes, big-endian x then y
    if (in_pos + 64 > in_len) return 1;
    uint8_t xb[32], yb[32];
    memcpy(xb, in + in_pos, 32); memcpy(yb, in + in_pos + 32, 32);
    in_pos += 64;
    reverse32(xb); reverse32(yb);
    if (!fq_from_repr(xb, &p.x) || !fq_from_repr(yb, &p.y)) return 1;
    Fq three = fe_zero<FqP>(); three.v[0] = 3; three = fe_from_canonical(three);
    if (!fe_eq(fe_sqr(p.y), fe_add(fe_mul(fe_sqr(p.x), p.x), three))) return 1;
  } else {
    if (in_pos + 32 > in_len) return 1;
    uint8_t xb[32];
    memcpy(xb, in + in_pos, 32);
    in_pos += 32;
    int sign;
    if (point_format == 0) { sign = xb[31] >> 7; xb[31] &= 0x7f; }
    else { if (xb[31] & 0x80) return 1; sign = (xb[31] >> 6) & 1; xb[31] &= 0x3f; }
    if (!fq_from_repr(xb, &p.x)) return 1;
    if (fe_is_zero(p.x) && sign == 0 && point_format == 0) return 1;   // the identity encoding
    if (!g1_decompress(p.x, sign, &p.y)) return 1;
  }
  if (common_point(p)) return 1;
  *out = p;
  return 0;
}
int Transcript::read_scalar(Fr* out) {
  if (in_pos + 32 > in_len) return 1;
  uint8_t b[32];
  memcpy(b, in + in_pos, 32);
  in_pos += 32;
  if (kind == 2) reverse32(b);
  Fr v; memcpy(v.v, b, 32);
  if (geq_mod<FrP>(v.v)) return 1;
  *out = fe_from_canonical(v);
  common_scalar(*out);
  return 0;
}

}  // namespace host
}  // namespace zkc

// Host hashes of the transcripts, exposed so that they can be pinned against published vectors (RFC 7693, Keccak team
// test vectors) and hashlib without a GPU.  kind 0: Blake2b-512 with a 16-byte personalisation (null = none), 64 bytes
// out; kind 1: Keccak-256 (pre-NIST padding), 32 bytes out.
extern "C" int zkc_host_hash(int kind, const uint8_t* personal16, const uint8_t* data, size_t len, uint8_t* out) {
  if (!out || (len && !data)) return 1;
  if (kind == 0) {
    char pers[16] = {0};
    if (personal16) memcpy(pers, personal16, 16);
    zkc::host::Blake2b b;
    b.init(pers);
    if (len) b.update(data, len);
    b.finalize(out);
    return 0;
  }
  if (kind == 1) { zkc::host::keccak256(data, len, out); return 0; }
  return 1;
}

// Host field inversion exposed for tests: which = 0 the product's host fe_inv (binary extended Euclid), 1 the Fermat ladder;
// field = 0 Fr, 1 Fq.  Montgomery in, Montgomery out.
extern "C" int zkc_host_fe_inv(int field, int which, const uint64_t* in, uint64_t* out, size_t n) {
  if (!in || !out || field < 0 || field > 1 || which < 0 || which > 1) return 1;
  for (size_t i = 0; i < n; ++i) {
    if (field == 0) {
      zkc::Fr a; memcpy(a.v, in + 4 * i, 32);
      if (zkc::geq_mod<zkc::FrP>(a.v)) return 1;
      const zkc::Fr r = which ? zkc::fe_inv_fermat_host(a) : zkc::fe_inv(a);
      memcpy(out + 4 * i, r.v, 32);
    } else {
      zkc::Fq a; memcpy(a.v, in + 4 * i, 32);
      if (zkc::geq_mod<zkc::FqP>(a.v)) return 1;
      const zkc::Fq r = which ? zkc::fe_inv_fermat_host(a) : zkc::fe_inv(a);
      memcpy(out + 4 * i, r.v, 32);
    }
  }
  return 0;
}
