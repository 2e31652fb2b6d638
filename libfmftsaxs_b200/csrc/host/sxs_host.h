/* Internal helpers shared by the host sources. */
#ifndef SXS_HOST_H
#define SXS_HOST_H
#include "common.h"
#include "sxs_cuda.h"

/* Device used by single-device entry points: first id of SXS_CUDA_DEVICES, else 0. */
int sxs_host_default_device(void);
/* Parses SXS_CUDA_DEVICES ("0,1,2"); default = all visible devices.  Returns the count. */
int sxs_host_device_list(int *devices, int cap);

#define SXS_CUDA_CHECK(call) do {                                                      \
	if ((call) != 0) {                                                                 \
		fprintf(stderr, "[Error] %s, function %s, line %i: CUDA layer: %s\n",          \
		        __FILE__, __func__, __LINE__, sxs_cuda_last_error());                  \
		exit(EXIT_FAILURE);                                                            \
	}                                                                                  \
} while (0)

#endif
