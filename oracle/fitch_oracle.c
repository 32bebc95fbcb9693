/* TEST INFRASTRUCTURE ONLY -- see oracle.h.  Per-pair Fitch operations, restated.
 *
 * The nibble loops follow the reference's readable byte-wise definition
 * (src/seq_b2_w1_basic.c:155-244: FitchBoth / FitchBases / FitchScoreWeight1 / FitchScore): for each
 * site, if the two state sets intersect keep the intersection, otherwise take the union and charge
 * the site's weight.  The stop-early / fw1 contract follows src/fitchtemplate_b2_w8_fastc64.inc
 * (:43-80 weight-1, :89-163 weighted): test after every block, hand over to the weight-1 loop at the
 * first block whose first weight is 1.  The SWAR forms restate src/seq_b2_w8_fastc64.c:11-23.
 */
#include "oracle.h"

static inline unsigned nib(const uint8_t *s, size_t site) {
	return (site & 1) ? (unsigned) (s[site >> 1] >> 4) : (unsigned) (s[site >> 1] & 0x0F);
}

void xo_fitch_bases(const uint8_t *s1, const uint8_t *s2, uint8_t *dest, size_t nbytes) {
	for (size_t b = 0; b < nbytes; ++b) {
		unsigned out = 0;
		for (unsigned h = 0; h < 2; ++h) {
			unsigned x = nib(s1, 2 * b + h), y = nib(s2, 2 * b + h);
			unsigned d = x & y;
			if (!d) d = x | y;
			out |= d << (4 * h);
		}
		dest[b] = (uint8_t) out;
	}
}

unsigned xo_fitch_bases_and_compare(const uint8_t *s1, const uint8_t *s2, uint8_t *dest, const uint8_t *prev, size_t nbytes) {
	unsigned changed = 0;
	xo_fitch_bases(s1, s2, dest, nbytes);
	for (size_t b = 0; b < nbytes; ++b) changed |= (dest[b] != prev[b]);
	return changed;
}

unsigned xo_fitch_score_weight1(const uint8_t *s1, const uint8_t *s2, size_t nbytes) {
	unsigned score = 0;
	for (size_t s = 0; s < 2 * nbytes; ++s) score += !(nib(s1, s) & nib(s2, s));
	return score;
}

unsigned xo_fitch_score(const uint8_t *s1, const uint8_t *s2, const unsigned *weights, size_t nbytes) {
	unsigned score = 0;
	for (size_t s = 0; s < 2 * nbytes; ++s) {
		if (!(nib(s1, s) & nib(s2, s))) score += weights[s];
	}
	return score;
}

unsigned xo_fitch_score_weight1_stopearly(const uint8_t *s1, const uint8_t *s2, size_t nbytes,
                                          unsigned block_bytes, unsigned max_score) {
	unsigned score = 0;
	for (size_t b0 = 0; b0 < nbytes; b0 += block_bytes) {
		for (size_t s = 2 * b0; s < 2 * (b0 + block_bytes); ++s) score += !(nib(s1, s) & nib(s2, s));
		if (score > max_score) return score;
	}
	return score;
}

unsigned xo_fitch_score_stopearly(const uint8_t *s1, const uint8_t *s2, const unsigned *weights, size_t nbytes,
                                  unsigned block_bytes, int fw1, unsigned max_score) {
	unsigned score = 0;
	for (size_t b0 = 0; b0 < nbytes; b0 += block_bytes) {
		if (fw1 && weights[2 * b0] == 1) {
			return score + xo_fitch_score_weight1_stopearly(s1 + b0, s2 + b0, nbytes - b0, block_bytes, max_score);
		}
		for (size_t s = 2 * b0; s < 2 * (b0 + block_bytes); ++s) {
			if (!(nib(s1, s) & nib(s2, s))) score += weights[s];
		}
		if (score > max_score) return score;
	}
	return score;
}

/* 64-bit SWAR: z has the top bit of every nibble set iff that nibble of (a & b) is non-empty. */
#define LO3 0x7777777777777777ULL
#define HI1 0x8888888888888888ULL
#define ONE 0x1111111111111111ULL

void xo_fitch_bases_swar64(const uint64_t *s1, const uint64_t *s2, uint64_t *dest, size_t nwords) {
	for (size_t i = 0; i < nwords; ++i) {
		uint64_t x = s1[i] & s2[i];
		uint64_t z = (((x & LO3) + LO3) | x) & HI1;       /* non-empty marker, bit 3 of each nibble */
		uint64_t keep = (z >> 3) * 15;                     /* 0xF where the intersection is non-empty */
		dest[i] = x | ((s1[i] | s2[i]) & ~keep);
	}
}

unsigned xo_fitch_score_weight1_swar64(const uint64_t *s1, const uint64_t *s2, size_t nwords) {
	unsigned score = 0;
	for (size_t i = 0; i < nwords; ++i) {
		uint64_t x = s1[i] & s2[i];
		uint64_t z = (((x & LO3) + LO3) | x) >> 3;
		score += (unsigned) __builtin_popcountll(~z & ONE);
	}
	return score;
}

/* Weighted cost, 64 bits at a time: visit only the sites whose sets do not intersect. */
unsigned xo_fitch_score_swar64(const uint64_t *s1, const uint64_t *s2, const unsigned *weights, size_t nwords) {
	unsigned score = 0;
	for (size_t i = 0; i < nwords; ++i) {
		uint64_t x = s1[i] & s2[i];
		uint64_t e = ~((((x & LO3) + LO3) | x) >> 3) & ONE;      /* bit 0 of each empty nibble */
		while (e) {
			score += weights[16 * i + ((unsigned) __builtin_ctzll(e) >> 2)];
			e &= e - 1;
		}
	}
	return score;
}
