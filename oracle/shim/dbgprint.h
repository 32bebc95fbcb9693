/* TEST INFRASTRUCTURE ONLY (oracle/_ref build glue).
 *
 * Stand-in for <dbgprint.h> from wtwhite/useful_libs, a third-party dependency
 * of the reference that is not vendored under /root/reference (see the
 * reference's README.md:15-29 and src/CMakeLists.txt:7-74).  The reference only
 * uses DBGPRINT1..DBGPRINT8(fmt, args...) which print to `dbgfile` in DEBUG
 * builds and vanish otherwise.  No path arithmetic lives in that library.
 */
#ifndef XMP_ORACLE_SHIM_DBGPRINT_H
#define XMP_ORACLE_SHIM_DBGPRINT_H
#include <stdio.h>
extern FILE *dbgfile;
#ifdef DEBUG
#define XMP_DBG(...) fprintf(dbgfile, __VA_ARGS__)
#else
#define XMP_DBG(...) ((void) 0)
#endif
#define DBGPRINT1(f) XMP_DBG(f)
#define DBGPRINT2(f, a) XMP_DBG(f, a)
#define DBGPRINT3(f, a, b) XMP_DBG(f, a, b)
#define DBGPRINT4(f, a, b, c) XMP_DBG(f, a, b, c)
#define DBGPRINT5(f, a, b, c, d) XMP_DBG(f, a, b, c, d)
#define DBGPRINT6(f, a, b, c, d, e) XMP_DBG(f, a, b, c, d, e)
#define DBGPRINT7(f, a, b, c, d, e, g) XMP_DBG(f, a, b, c, d, e, g)
#define DBGPRINT8(f, a, b, c, d, e, g, h) XMP_DBG(f, a, b, c, d, e, g, h)
#endif
