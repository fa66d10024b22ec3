// Source-compatibility header (worldb200): the linkage macros the reference's public headers use
// (/root/reference/include/macrodefinitions.hpp).  Only the extern "C" pair is meaningful here.
#ifndef WORLD_MACRODEFINITIONS_HPP
#define WORLD_MACRODEFINITIONS_HPP

#ifdef __cplusplus
#define WORLD_BEGIN_C_DECLS extern "C" {
#define WORLD_END_C_DECLS }
#else
#define WORLD_BEGIN_C_DECLS
#define WORLD_END_C_DECLS
#endif

#ifndef WORLD_API
#define WORLD_API
#endif

#endif
