/* PHYLIP alignment reader -- behaviour of the reference's ReadPhylipAlignment
 * (reference src/common.c:2850-2948), line-oriented instead of character-oriented.
 *
 * Accepted input, as in the reference: a "<taxa> <sites>" header line; sequential or interleaved
 * blocks; in the first block the first 10 characters of every non-blank line are the label (trailing
 * blanks trimmed, blanks inside kept); sequence characters are IUPAC codes, '-', '?' in either case;
 * white space and digits between them are ignored; blank lines are skipped anywhere; the last line
 * may lack its newline.  Every line of a block must carry as many sites as the block's first line.
 */
#include <ctype.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "xmp_host.h"

static char g_err[512];

const char *xmph_last_error(void) { return g_err; }

/* Shared by the other host modules. */
int xmph_fail(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	return -1;
}

typedef struct {
	FILE *f;
	char *buf;
	size_t len, cap;
	int hit_eof;
} line_reader;

/* Reads up to and excluding the next '\n'.  Sets hit_eof once the stream has reported end of file. */
static void next_line(line_reader *r) {
	int c;
	r->len = 0;
	while ((c = getc_unlocked(r->f)) != EOF && c != '\n') {      /* one reader per stream: no need to lock it per character */
		if (r->len + 1 >= r->cap) {
			r->cap = r->cap ? r->cap * 2 : 256;
			r->buf = (char *) realloc(r->buf, r->cap);
		}
		r->buf[r->len++] = (char) c;
	}
	if (c == EOF) r->hit_eof = 1;
}

static int is_blank(const char *s, size_t n) {
	for (size_t i = 0; i < n; ++i) {
		if (!isspace((unsigned char) s[i])) return 0;
	}
	return 1;
}

void xmph_free_alignment(xmph_alignment *a) {
	if (a->labels) {
		for (unsigned i = 0; i < a->num_taxa; ++i) free(a->labels[i]);
	}
	free(a->labels);
	free(a->chars);
	memset(a, 0, sizeof *a);
}

int xmph_read_phylip(FILE *in, xmph_alignment *a) {
	static const char allowed[] = "ACGTURYSWKMBDHV-N?";
	line_reader r = { in, NULL, 0, 0, 0 };
	unsigned done = 0;          /* sites completed in earlier blocks */
	unsigned line_no = 2;
	int rc = 0;

	/* character classes: the (upper-cased) state symbol, IGNORED for white space and digits, 0 for anything else */
	enum { IGNORED = 0xFF };
	unsigned char cls[256];
	for (int v = 0; v < 256; ++v) {
		const int u = toupper(v);
		cls[v] = (unsigned char) ((u && strchr(allowed, u)) ? u : (isspace(u) || isdigit(u)) ? IGNORED : 0);
	}

	memset(a, 0, sizeof *a);
	next_line(&r);
	if (r.hit_eof && r.len == 0) {
		free(r.buf);
		return xmph_fail("Error reading input file: no first line!  Aborting.");
	}
	r.buf = (char *) realloc(r.buf, r.len + 1);
	r.cap = r.len + 1;
	r.buf[r.len] = 0;
	if (sscanf(r.buf, "%u %u", &a->num_taxa, &a->seq_len) != 2 || a->num_taxa == 0) {
		free(r.buf);
		return xmph_fail("Error reading input file: first line must give the number of taxa and sites, aborting.");
	}
	a->chars = (unsigned char *) malloc((size_t) a->num_taxa * a->seq_len + 1);
	a->labels = (char **) calloc(a->num_taxa, sizeof(char *));

	while (done < a->seq_len && !rc) {
		unsigned width = 0;     /* sites on the first line of this block */
		unsigned got = 0;
		for (unsigned t = 0; t < a->num_taxa && !rc; ++line_no) {
			if (r.hit_eof) {
				if (t) {
					rc = xmph_fail("Error reading input file: %u sites specified in top line, but input only contains %u sites for first %u taxa and %u sites for remaining %u taxa, aborting.",
					               a->seq_len, done + width, t, done, a->num_taxa - t);
				} else {
					rc = xmph_fail("Error reading input file: %u sites specified in top line, but input only contains %u sites, aborting.", a->seq_len, done);
				}
				break;
			}
			next_line(&r);
			if (is_blank(r.buf, r.len)) continue;

			size_t start = 0;
			if (done == 0) {
				/* Label: the first 10 characters, trailing blanks dropped. */
				size_t n = r.len < 10 ? r.len : 10;
				while (n > 0 && isspace((unsigned char) r.buf[n - 1])) --n;
				a->labels[t] = (char *) malloc(n + 1);
				memcpy(a->labels[t], r.buf, n);
				a->labels[t][n] = 0;
				start = r.len < 10 ? r.len : 10;
			}
			got = 0;
			unsigned char *dst = a->chars + (size_t) t * a->seq_len + done;
			const unsigned room = a->seq_len - done;
			for (size_t i = start; i < r.len && !rc; ++i) {
				const unsigned char c = cls[(unsigned char) r.buf[i]];
				if (c == IGNORED) continue;
				if (c == 0) {
					rc = xmph_fail("Error reading input file: invalid character '%c' on line %u, aborting.", r.buf[i], line_no);
				} else if (got == room) {
					rc = xmph_fail("Error reading input file: too many characters on line %u, aborting.", line_no);
				} else {
					dst[got++] = c;
				}
			}
			if (rc) break;
			if (t == 0) width = got;
			if (got != width) {
				rc = xmph_fail("Error reading input file: line %u contains %u %s sites than the first line in this group, aborting.",
				               line_no, got > width ? got - width : width - got, got > width ? "more" : "fewer");
				break;
			}
			++t;
		}
		done += got;
		if (!rc && got == 0) {
			rc = xmph_fail("Error reading input file: %u sites specified in top line, but input only contains %u sites, aborting.", a->seq_len, done);
		}
	}
	free(r.buf);
	if (rc) xmph_free_alignment(a);
	return rc;
}
