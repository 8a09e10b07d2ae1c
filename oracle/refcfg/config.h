/* Hand-written stand-in for the config.h that the reference's ./configure would
 * generate.  Test infrastructure only (see oracle/README.md): it lets the recipe in
 * oracle/Makefile compile the reference's jpeg2000 sources *where they lie* under
 * /root/reference with plain gcc, without running the reference's build system.
 * Generic C, no assembly, no optional libraries: the same choices the survey's
 * --disable-x86asm --disable-everything scratch build made. */
#ifndef ORACLE_REFCFG_CONFIG_H
#define ORACLE_REFCFG_CONFIG_H

#define FFMPEG_CONFIGURATION "oracle-refcfg"
#define FFMPEG_LICENSE "LGPL version 2.1 or later"
#define av_restrict restrict
#define EXTERN_PREFIX ""
#define EXTERN_ASM
#define BUILDSUF ""
#define SLIBSUF ".so"
#define SWS_MAX_FILTER_SIZE 256

/* architectures: none selected => portable C paths only */
#define ARCH_AARCH64 0
#define ARCH_ALPHA 0
#define ARCH_ARM 0
#define ARCH_LOONGARCH 0
#define ARCH_LOONGARCH32 0
#define ARCH_LOONGARCH64 0
#define ARCH_MIPS 0
#define ARCH_MIPS64 0
#define ARCH_PPC 0
#define ARCH_PPC64 0
#define ARCH_RISCV 0
#define ARCH_S390 0
#define ARCH_SPARC 0
#define ARCH_X86 0
#define ARCH_X86_32 0
#define ARCH_X86_64 0

#define HAVE_BIGENDIAN 0
#define HAVE_FAST_UNALIGNED 1
#define HAVE_FAST_CLZ 1
#define HAVE_FAST_64BIT 1
#define HAVE_FAST_CMOV 0
#define HAVE_INLINE_ASM 0
#define HAVE_MMX 0
#define HAVE_MMX_INLINE 0
#define HAVE_MMX_EXTERNAL 0
#define HAVE_MM_EMPTY 0
#define HAVE_MIPSFPU 0
#define HAVE_SIMD_ALIGN_16 0
#define HAVE_SIMD_ALIGN_32 0
#define HAVE_SIMD_ALIGN_64 0
#define HAVE_PRAGMA_DEPRECATED 1
#define HAVE_LIBC_MSVCRT 0

#define HAVE_THREADS 1
#define HAVE_PTHREADS 1
#define HAVE_W32THREADS 0
#define HAVE_OS2THREADS 0
#define HAVE_PTHREAD_NP_H 0
#define HAVE_PTHREAD_SETNAME_NP 0
#define HAVE_PTHREAD_SET_NAME_NP 0
#define HAVE_PRCTL 0
#define HAVE_CLOCK_GETTIME 1
#define HAVE_GETHRTIME 0
#define HAVE_MACH_ABSOLUTE_TIME 0

/* libm: everything is present in glibc */
#define HAVE_ATANF 1
#define HAVE_ATAN2F 1
#define HAVE_CBRT 1
#define HAVE_CBRTF 1
#define HAVE_COPYSIGN 1
#define HAVE_COSF 1
#define HAVE_ERF 1
#define HAVE_EXP2 1
#define HAVE_EXP2F 1
#define HAVE_EXPF 1
#define HAVE_HYPOT 1
#define HAVE_ISFINITE 1
#define HAVE_ISINF 1
#define HAVE_ISNAN 1
#define HAVE_LDEXPF 1
#define HAVE_LLRINT 1
#define HAVE_LLRINTF 1
#define HAVE_LOG2 1
#define HAVE_LOG2F 1
#define HAVE_LOG10F 1
#define HAVE_LRINT 1
#define HAVE_LRINTF 1
#define HAVE_POWF 1
#define HAVE_RINT 1
#define HAVE_ROUND 1
#define HAVE_ROUNDF 1
#define HAVE_SINF 1
#define HAVE_TRUNC 1
#define HAVE_TRUNCF 1

#define CONFIG_SMALL 0
#define CONFIG_SHARED 0
#define CONFIG_LCMS2 0
#define CONFIG_LINUX_PERF 0
#define CONFIG_MACOS_KPERF 0
#define CONFIG_COLOR_RANGE 0
#define CONFIG_MEMORY_POISONING 0
#define CONFIG_FRAME_THREAD_ENCODER 0
#define CONFIG_GRAY 0
#define CONFIG_ICONV 0

#endif
