/* lol_internal.h -- helpers shared by the C translation units of liblolb200. */
#ifndef LOL_INTERNAL_H
#define LOL_INTERNAL_H

#include "lolb200.h"

#ifdef __cplusplus
extern "C" {
#endif

void lolb200_set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
int lolb200_scene_check(const lolb200_scene* s);
/* Scene-dependent licences for the exact work-skipping shortcuts (lol_lower.c). */
int lolb200_can_skip_black_miss(const lolb200_scene* s);
int lolb200_can_cull_backfacing(const lolb200_scene* s);
int lolb200_can_shadow_early(const lolb200_scene* s);
/* 1 when n*rk corrected by two FMAs equals n/k for every significand of n. */
int lolb200_div_const_is_exact(float k);

/* Threads per CTA of the generated kernel (8 warps). */
#define LOLB200_KERNEL_THREADS 256
/* Kernel structure used when options.variant == 0 (1 phase-sequential, 2 compaction). */
#define LOLB200_DEFAULT_VARIANT 1
/* options.roll_v1 = -1: what variants 1 and 4 roll into loops by default (0 nothing, 1 normal taps, 2 + lights) */
#define LOLB200_DEFAULT_ROLL_V1 0

#ifdef __cplusplus
}
#endif
#endif
