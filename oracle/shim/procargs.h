/* TEST INFRASTRUCTURE ONLY (oracle/_ref build glue).
 *
 * Stand-in for <procargs.h> from wtwhite/useful_libs (un-vendored, unpinned).
 * Call sites: reference src/common.c:179-197.  The "--name={Y|N}" and
 * "--name=<nn>" syntax is taken from PrintUsage (src/common.c:663-670) and the
 * author's perf scripts (tools/orac_perf_test.pl:40).
 */
#ifndef XMP_ORACLE_SHIM_PROCARGS_H
#define XMP_ORACLE_SHIM_PROCARGS_H
int ProcessBooleanExtOption(char *arg, char *name, unsigned *flags, unsigned mask);
int ProcessUnsignedExtOption(char *arg, char *name, unsigned *out);
#endif
