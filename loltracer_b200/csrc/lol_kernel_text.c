/*
 * lol_kernel_text.c -- embeds lol_kernel.cuh and lol_params.h into the library as a C string.
 * The build passes -DLOL_KERNEL_CUH="/abs/path/lol_kernel.cuh"
 * and -DLOL_PARAMS_H_PATH="/abs/path/lol_params.h".
 */
#if !defined(LOL_KERNEL_CUH) || !defined(LOL_PARAMS_H_PATH)
#error "define LOL_KERNEL_CUH and LOL_PARAMS_H_PATH to the paths of lol_kernel.cuh / lol_params.h"
#endif

__asm__(".section .rodata\n"
        ".global lol_kernel_text\n"
        ".type lol_kernel_text, @object\n"
        "lol_kernel_text:\n"
        ".incbin \"" LOL_KERNEL_CUH "\"\n"
        ".byte 0\n"
        ".size lol_kernel_text, .-lol_kernel_text\n"
        ".global lol_params_text\n"
        ".type lol_params_text, @object\n"
        "lol_params_text:\n"
        ".incbin \"" LOL_PARAMS_H_PATH "\"\n"
        ".byte 0\n"
        ".size lol_params_text, .-lol_params_text\n"
        ".previous\n");
