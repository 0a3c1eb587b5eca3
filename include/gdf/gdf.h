/* C / C++ umbrella for the gdf_* hot-path ABI (mirrors the role of reference libgdf/include/gdf/gdf.h:1-17). */
#ifndef GDF_B200_GDF_H
#define GDF_B200_GDF_H
#include <stddef.h>
#include <stdint.h>
#include "cffi/types.h"
#define GDF_VALID_BITSIZE 8 /* bits per gdf_valid_type, ref gdf.h:10 */
#ifdef __cplusplus
extern "C" {
#endif
#include "cffi/functions.h"
#ifdef __cplusplus
}
#endif
#endif
