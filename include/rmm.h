/* C / C++ wrapper around memory.h (role of reference libgdf/include/rmm.h). */
#ifndef GDF_B200_RMM_H
#define GDF_B200_RMM_H
#include <stddef.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif
#ifdef __cplusplus
extern "C" {
#endif
#include "memory.h"
#ifdef __cplusplus
}
#endif
#endif
