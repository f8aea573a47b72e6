/* Out-of-tree replacement for the header CMake would generate from
 * /root/reference/source/lib/PccLibBitstreamCommon/include/PCCConfig.h.in (the reference tree is
 * read-only, PccLibCommon/CMakeLists.txt:12-13 configures it in place).  No USE_*_VIDEO_CODEC macro is
 * defined: the oracle never touches the video codecs. */
#ifndef PCC_CONFIG_H_RABBIT_B200_ORACLE
#define PCC_CONFIG_H_RABBIT_B200_ORACLE
#define TMC2_VERSION_MAJOR 15
#define TMC2_VERSION_MINOR 0
#define HAVE_GETRUSAGE 1
#endif
