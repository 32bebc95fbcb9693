/* seq_b2_w16_cuda.h -- the B200 Fitch family as the reference's seq.h sees it.
 *
 * A family header of the reference's own kind (compare src/seq_b2_sse2asm.h:13-32 and src/seq_b2_w8_fastc64.h:7-21):
 * it declares the family's functions and, when the family is the selected one (CUDAFITCH, next to BASICFITCH /
 * FASTCFITCH / FASTC64FITCH / SSE2ASMFITCH in switches.h), nominates them under the suffix-less `_b2_w16` names that
 * the CHOSEN_* dispatch of src/seq.h:29-66 pastes together.  With SITESPERBYTE 2 and BYTESPERBLOCK 16 every
 * CHOSEN_FitchBases / CHOSEN_FitchBasesAndCompare / CHOSEN_FitchScore / CHOSEN_FitchScoreWeight1 call of the
 * reference (src/yanbader.c:160-166,336-342,364-369,420-425, src/common.c:569-610, src/greedytree.c:28-34, ...) then
 * lands in libxmp_b200.so.
 *
 * What a maintainer adds to the reference (INTEGRATION.md shows it; this repository's test build does exactly this at
 * build time, on a generated copy of seq.h, to prove that the reference links and runs against the library -- the
 * `ref_cuda` target described there):
 *     switches.h.in   #cmakedefine CUDAFITCH                                  (one line)
 *     seq.h:5         ... + defined(SSE2ASMFITCH) + defined(CUDAFITCH) != 1   (the one-family check)
 *     seq.h:20        #include "seq_b2_w16_cuda.h"                            (after seq_b2_sse2asm.h)
 *     CMakeLists.txt  IF(CUDAFITCH) SET(BYTESPERBLOCK 16 ...) + link xmp_b200 (like the SSE2ASMFITCH block, :264-270)
 *
 * One launch per call is never the fast path -- the batched surface of xmp_cuda.h is -- but it is the reference's
 * own call surface, bit for bit: same arguments, same results (the stop-early variants return the exact score,
 * which their contract allows).
 */
#ifndef SEQ_B2_W16_CUDA_H
#define SEQ_B2_W16_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

/* There is no uint128_t type: void *, as in src/seq_b2_sse2asm.h:13-14.  Buffers are HOST memory. */
void     FitchBases_b2_w16_cuda(void *s1, void *s2, void *dest, unsigned nBlocks);
unsigned FitchBasesAndCompare_b2_w16_cuda(void *s1, void *s2, void *dest, void *prev, unsigned nBlocks);
unsigned FitchScoreWeight1_b2_w16_cuda(void *s1, void *s2, unsigned nBlocks);
unsigned FitchScoreWeight1_stopearly_b2_w16_cuda(void *s1, void *s2, unsigned nBlocks, unsigned maxScore);
unsigned FitchScore_b2_w16_cuda(void *s1, void *s2, unsigned *weights, unsigned nBlocks);
unsigned FitchScore_stopearly_b2_w16_cuda(void *s1, void *s2, unsigned *weights, unsigned nBlocks, unsigned maxScore);
unsigned FitchScore_fw1_b2_w16_cuda(void *s1, void *s2, unsigned *weights, unsigned nBlocks);
unsigned FitchScore_fw1_stopearly_b2_w16_cuda(void *s1, void *s2, unsigned *weights, unsigned nBlocks, unsigned maxScore);

#ifdef __cplusplus
}
#endif

/* If we have been selected as the implementation, export macros for the briefer names (src/seq_b2_sse2asm.h:23-32). */
#ifdef CUDAFITCH
#define FitchBases_b2_w16 FitchBases_b2_w16_cuda
#define FitchBasesAndCompare_b2_w16 FitchBasesAndCompare_b2_w16_cuda
#define FitchScoreWeight1_b2_w16 FitchScoreWeight1_b2_w16_cuda
#define FitchScoreWeight1_stopearly_b2_w16 FitchScoreWeight1_stopearly_b2_w16_cuda
#define FitchScore_b2_w16 FitchScore_b2_w16_cuda
#define FitchScore_stopearly_b2_w16 FitchScore_stopearly_b2_w16_cuda
#define FitchScore_fw1_b2_w16 FitchScore_fw1_b2_w16_cuda
#define FitchScore_fw1_stopearly_b2_w16 FitchScore_fw1_stopearly_b2_w16_cuda
#endif /* CUDAFITCH */
#endif
