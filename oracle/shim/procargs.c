/* TEST INFRASTRUCTURE ONLY (oracle/_ref build glue); see procargs.h. */
#include <stdlib.h>
#include <string.h>
#include "procargs.h"

/* Returns the text after "name" or "name=" if arg names this option, else NULL. */
static const char *match_option(const char *arg, const char *name) {
	size_t n = strlen(name);
	if (strncmp(arg, name, n) != 0) return NULL;
	if (arg[n] == 0) return arg + n;
	if (arg[n] == '=') return arg + n + 1;
	return NULL;
}

int ProcessBooleanExtOption(char *arg, char *name, unsigned *flags, unsigned mask) {
	const char *v = match_option(arg, name);
	if (!v) return 0;
	if (*v == 0 || *v == 'Y' || *v == 'y' || *v == '1' || *v == 'T' || *v == 't') {
		*flags |= mask;
	} else {
		*flags &= ~mask;
	}
	return 1;
}

int ProcessUnsignedExtOption(char *arg, char *name, unsigned *out) {
	const char *v = match_option(arg, name);
	if (!v || *v == 0) return 0;
	*out = (unsigned) strtoul(v, NULL, 10);
	return 1;
}
