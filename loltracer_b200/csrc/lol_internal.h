/* lol_internal.h -- helpers shared by the C translation units of liblolb200. */
#ifndef LOL_INTERNAL_H
#define LOL_INTERNAL_H

#include "lolb200.h"

#ifdef __cplusplus
extern "C" {
#endif

void lolb200_set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
int lolb200_scene_check(const lolb200_scene* s);

#ifdef __cplusplus
}
#endif
#endif
