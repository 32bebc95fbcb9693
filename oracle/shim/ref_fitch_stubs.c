/* TEST INFRASTRUCTURE ONLY (oracle/_ref build glue).
 *
 * Link-time stubs so that the reference's three Fitch source files
 * (seq_b2_w8_fastc64.c, seq_b2_w4_fastc.c, seq_b2_w1_basic.c) can be built into
 * a stand-alone shared library without the rest of the program.  Only globals
 * and one helper that those files reference are provided; no arithmetic.
 */
#include <stdio.h>
#include <stdlib.h>
#include "common.h"

FILE *dbgfile;
struct CharData gCD;
struct TreeData gTD;
unsigned _ColumnCompareWidth;

char GetBaseFromMask(unsigned char mask) {
	static const char names[16] = { '0', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N' };
	return names[mask & 15];
}
