/* Shared includes, error macros and the (l, m) -> flat index map; mirrors src/common.h. */
#ifndef FMFTSAXS_COMMON_H
#define FMFTSAXS_COMMON_H

#include <assert.h>
#include <ctype.h>
#include <errno.h>
#include <getopt.h>
#include <libgen.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "define.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Errors at the public boundary print and terminate the process, like src/common.h:26-36. */
#define ERROR_MSG(msg) do {                                                              \
	fprintf(stderr, "[Error] %s, function %s, line %i: %s\n", __FILE__, __func__, __LINE__, msg); \
	exit(EXIT_FAILURE);                                                                  \
} while (0)

#define CHECK_PTR(ptr) do { if ((ptr) == NULL) { ERROR_MSG("Null pointer"); } } while (0)

#ifdef _SXS_VERBOSE_
#define SXS_PRINTF(...) do { printf(__VA_ARGS__); } while (0)
#else
#define SXS_PRINTF(...) do { } while (0)
#endif

/* A_{lm} lives at l(l+1)+m (src/common.h:49-51). */
static inline int lm_index(int l, int m) { return l * (l + 1) + m; }

static inline void sxs_myfree(void *ptr) { if (ptr != NULL) { free(ptr); } }

static inline FILE *sxs_myfopen(char *path, char *mode)
{
	FILE *f = fopen(path, mode);
	if (f == NULL) {
		fprintf(stderr, "Unable to open %s\n", path);
		exit(EXIT_FAILURE);
	}
	return f;
}

#ifdef __cplusplus
}
#endif

#endif
