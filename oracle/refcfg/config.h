/* Hand-written stand-in for the config.h meson would generate for the
 * reference's C sources (portable C path, no assembly).  Test infrastructure
 * only: used by oracle/Makefile to compile /root/reference/src in place. */
#ifndef RAV1D_B200_ORACLE_CONFIG_H
#define RAV1D_B200_ORACLE_CONFIG_H
#define ARCH_AARCH64 0
#define ARCH_ARM 0
#define ARCH_PPC64LE 0
#define ARCH_X86 0
#define ARCH_X86_32 0
#define ARCH_X86_64 0
#define HAVE_ASM 0
#define CONFIG_8BPC 1
#define CONFIG_16BPC 1
#define CONFIG_LOG 1
#define ENDIANNESS_BIG 0
#define HAVE_POSIX_MEMALIGN 1
#define HAVE_UNISTD_H 1
#define HAVE_CLOCK_GETTIME 1
#define HAVE_DLSYM 1
#define STACK_ALIGNMENT 16
#define TRIM_DSP_FUNCTIONS 0
#endif
