// Host-side pieces of the create_proof driver that stay on the CPU in the reference too:
// Fiat-Shamir transcripts (halo2_proofs::transcript::{Blake2bWrite, Keccak256Write}, Challenge255),
// the ChaCha20 RNG that feeds Fr::random (rand_chacha 0.3.1 / halo2curves `from_u512`), and small
// scalar helpers.  SURVEY.md §8a rows a13/a14, Appendix A.2/A.12.  Upstream sources are not vendored
// (/root/reference/Cargo.lock:313 blake2b_simd, :2635 sha3, :2195 rand_chacha, :1320 halo2_proofs).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "../ff.cuh"
#include "../ec.cuh"

namespace zkc {
namespace host {

// ---- Blake2b-512 with personalisation (RFC 7693) ----------------------------------------------------
struct Blake2b {
  uint64_t h[8];
  uint64_t t[2];
  uint8_t buf[128];
  size_t buflen;
  void init(const char personal[16]);
  void update(const uint8_t* in, size_t len);
  void finalize(uint8_t out[64]) const;   // does not disturb the running state (clone + finalize)
 private:
  void compress(const uint8_t block[128], bool last);
};

// ---- Keccak-256 (original padding 0x01) ---------------------------------------------------------------
void keccak256(const uint8_t* in, size_t len, uint8_t out[32]);

// ---- ChaCha20 RNG as rand_chacha::ChaCha20Rng (64-bit block counter, stream 0) ----------------------
struct ChaCha20Rng {
  uint32_t key[8];
  uint64_t counter;
  uint32_t block[16];
  int pos;  // next unread word in block; 16 = empty
  int double_rounds;   // 10 = ChaCha20Rng, 6 = ChaCha12Rng (rand 0.8 StdRng)
  explicit ChaCha20Rng(const uint8_t seed[32], int double_rounds = 10);
  uint32_t next_u32();
  uint64_t next_u64();
  Fr fr_random();   // halo2curves Fr::random: 8 x next_u64 as a 512-bit LE integer, reduced mod r
  // The generator is a linear stream of 32-bit keystream words (rand_core BlockRng): next_u64 takes two, Fr::random
  // sixteen, fill_bytes(32) eight.  word_pos() is the index of the next unread word; seek() moves there.
  uint64_t word_pos() const { return counter * 16 - (uint64_t)(16 - pos); }
  void seek(uint64_t word);
  void fill_bytes(uint8_t* out, size_t nbytes);   // nbytes a multiple of 4 (whole words)
};
// rand_core SeedableRng::seed_from_u64 (PCG32 expansion)
void seed_from_u64(uint64_t state, uint8_t seed[32]);

// ---- scalar helpers ---------------------------------------------------------------------------------------
Fr fr_from_u512_le(const uint8_t bytes[64]);      // from_uniform_bytes
void fr_to_repr(const Fr& a, uint8_t out[32]);    // canonical little-endian
void fq_to_repr(const Fq& a, uint8_t out[32]);
int fr_cmp_canonical(const Fr& a, const Fr& b);   // Ord for Fr

// ---- Poseidon sponge of snark-verifier's PoseidonTranscript (T = 3, RATE = 2, R_F = 8, R_P = 57; SURVEY OPEN-7) ----
struct PoseidonSponge {
  Fr state[3];
  std::vector<Fr> buf;
  PoseidonSponge();
  void update(const Fr* e, size_t n) { buf.insert(buf.end(), e, e + n); }
  Fr squeeze();
 private:
  void absorb(const Fr* chunk, size_t len);
};
// round constants [65][3] and MDS [3][3] from the Grain LFSR (Montgomery form); exposed for tests
void poseidon_spec(const Fr** constants, const Fr** mds);

// ---- transcripts --------------------------------------------------------------------------------------------
struct Transcript {
  int kind;          // 0 = Blake2bWrite, 1 = Keccak256Write (halo2), 2 = snark-verifier EvmTranscript (Keccak, big-endian,
                     // uncompressed points), 3 = snark-verifier PoseidonTranscript (what gen_snark_shplonk instantiates)
  PoseidonSponge pos;
  int point_format;  // SURVEY OPEN-5
  Blake2b b2;
  std::vector<uint8_t> kbuf;
  std::vector<uint8_t> proof;
  Transcript(int kind, int point_format);
  void absorb(const uint8_t* p, size_t n);
  Fr squeeze_challenge();
  int common_point(const G1Affine& p);   // returns nonzero on identity
  void common_scalar(const Fr& s);
  int write_point(const G1Affine& p);
  void write_scalar(const Fr& s);
  // reader side (TranscriptRead of the same four kinds): elements are taken from `in`, absorbed, and returned
  const uint8_t* in = nullptr;
  size_t in_len = 0, in_pos = 0;
  void set_input(const uint8_t* p, size_t n) { in = p; in_len = n; in_pos = 0; }
  int read_point(G1Affine* out);   // nonzero: truncated input, non-canonical coordinate, not on the curve, or identity
  int read_scalar(Fr* out);        // nonzero: truncated input or non-canonical scalar
};
// y with y^2 = x^3 + 3 and the given parity; false if x^3 + 3 is not a square
bool g1_decompress(const Fq& x, int y_is_odd, Fq* y);

}  // namespace host
}  // namespace zkc
