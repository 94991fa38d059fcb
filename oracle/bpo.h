/* oracle/bpo.h -- C CPU oracle for the Bulletproofs R1CS hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or load this library.  The product (libbpg.so) never does.
 *
 * Restates (reference call sites; the arithmetic itself lives in un-vendored crates):
 *   curve25519-dalek 1.x  Scalar / RistrettoPoint / Straus / Pippenger   (Cargo.toml:8)
 *   merlin 1.x            STROBE-128 transcript                          (Cargo.toml:10)
 *   bulletproofs develop  r1cs::Prover::prove, Verifier::verify, IPP     (Cargo.toml:17-20)
 *   in-tree               src/mimc_hash/mimc.rs, merkle_tree_gadget.rs:106
 * Parity status: MiMC pinned by the reference's KATs; group arithmetic pinned by RFC 9496 and
 * libsodium; MSM / Pedersen / IPP / proof bytes: "parity unpinned" (no golden bytes upstream).
 * All scalars: 32-byte little endian.  All points: 32-byte compressed ristretto255.
 */
#ifndef BPO_H
#define BPO_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* variable encoding inside constraint terms: kind << 29 | index */
#define BPO_VAR_L 0u
#define BPO_VAR_R 1u
#define BPO_VAR_O 2u
#define BPO_VAR_V 3u
#define BPO_VAR_ONE 4u

/* MSM algorithms */
#define BPO_MSM_NAIVE 0   /* double-and-add per point, the "obviously right" check      */
#define BPO_MSM_STRAUS_CT 1 /* dalek multiscalar_mul: radix-16 Straus                   */
#define BPO_MSM_VARTIME 2 /* dalek vartime_multiscalar_mul: NAF Straus <190, else Pippenger */

void bpo_set_mimc_constants(const uint8_t *consts486x32);
void bpo_set_threads(int n); /* OpenMP threads for the data-parallel loops (default 1)   */

/* scalar helpers (Scalar semantics) */
void bpo_sc_reduce(const uint8_t in32[32], uint8_t out32[32]);
void bpo_sc_wide(const uint8_t in64[64], uint8_t out32[32]);
void bpo_sc_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);
void bpo_sc_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);
void bpo_sc_invert(const uint8_t a[32], uint8_t out[32]);

/* group */
int bpo_point_decode_ok(const uint8_t p32[32]);
int bpo_point_add(const uint8_t a32[32], const uint8_t b32[32], uint8_t out32[32]);
int bpo_point_mul(const uint8_t s32[32], const uint8_t p32[32], uint8_t out32[32]);
void bpo_from_uniform_bytes(const uint8_t in64[64], uint8_t out32[32]);
void bpo_pedersen_gens(uint8_t B32[32], uint8_t Bblind32[32]);
void bpo_gens(size_t i0, size_t n, uint8_t *G32, uint8_t *H32); /* BulletproofGens party 0 */
void bpo_pedersen_commit(const uint8_t *v, const uint8_t *r, size_t n, uint8_t *out32);
int bpo_msm(const uint8_t *scalars, const uint8_t *points32, size_t n, uint8_t out32[32], int algo);
/* sum sG[i]*G[off+i] + sH[i]*H[off+i] (+ extra) over the BulletproofGens chains */
int bpo_msm_gens(const uint8_t *sG, const uint8_t *sH, size_t n, size_t offset,
                 const uint8_t *extra_scalars, const uint8_t *extra_points32, size_t k,
                 uint8_t out32[32], int algo);
/* one IPP generator fold: out[i] = sl*PL[i] + sr*PR[i]  (dalek: 2-point vartime MSM per i) */
int bpo_fold_points(const uint8_t sl[32], const uint8_t sr[32], const uint8_t *PL32,
                    const uint8_t *PR32, size_t n, uint8_t *out32);

/* merlin */
typedef struct bpo_transcript bpo_transcript;
bpo_transcript *bpo_transcript_new(const uint8_t *label, size_t len);
void bpo_transcript_free(bpo_transcript *);
void bpo_transcript_append(bpo_transcript *, const uint8_t *label, size_t ll, const uint8_t *msg, size_t ml);
void bpo_transcript_challenge(bpo_transcript *, const uint8_t *label, size_t ll, uint8_t *out, size_t n);

/* MiMC (src/mimc_hash/mimc.rs) */
void bpo_mimc_hash(const uint8_t *data, size_t len, uint8_t out32_le[32]);
/* unpadded sponge over nblocks 32-byte LE scalars; optional trace (972*3*32 B per block: per
   multiplier (aL,aR,aO) in gadget order mimc_hash_gadget.rs:133-144) */
void bpo_mimc_sponge(const uint8_t *blocks, size_t nblocks, uint8_t out32[32], uint8_t *trace);

/* R1CS: one call = Prover::new + commit* + (constraints given flat) + prove.
 * Transcript label = `label`.  V commitments are computed from v/v_blinding and appended in
 * order.  ext_rng32 = the 32 bytes dalek would draw from thread_rng in finalize().
 * Constraints in CSR: row_ptr[q+1], term_var[T], term_coeff[T*32].
 * Returns proof length (>0) or <0 on error.  proof_cap must be >= 1+32*(13+2*32+2).
 * flags bit0: legacy framing (no tag byte, 14 fixed fields). */
long bpo_r1cs_prove(const uint8_t *label, size_t label_len, size_t gens_capacity,
                    size_t n, const uint8_t *aL, const uint8_t *aR, const uint8_t *aO,
                    size_t m, const uint8_t *v, const uint8_t *v_blinding,
                    size_t q, const uint32_t *row_ptr, const uint32_t *term_var, const uint8_t *term_coeff,
                    const uint8_t ext_rng32[32], int flags,
                    uint8_t *V_out /* m*32, nullable */, uint8_t *proof, size_t proof_cap);
/* returns 1 accept, 0 reject (VerificationError / FormatError) */
int bpo_r1cs_verify(const uint8_t *label, size_t label_len, size_t gens_capacity,
                    size_t n, size_t m, const uint8_t *V32,
                    size_t q, const uint32_t *row_ptr, const uint32_t *term_var, const uint8_t *term_coeff,
                    const uint8_t *proof, size_t proof_len, const uint8_t ext_rng32[32], int flags);

#ifdef __cplusplus
}
#endif
#endif
