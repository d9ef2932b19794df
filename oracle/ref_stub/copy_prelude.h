// oracle/ref_stub/copy_prelude.h -- TEST INFRASTRUCTURE ONLY.
// Minimal Windows-type prelude so that the sample->texture repack functions of the reference
// (Source/Helper.cpp:414-803, the CopyFrame* family) compile where they lie: the Makefile pipes
// this file, the reference's own lines (sed -n, nothing is copied into the repo) and
// ref_stub/copy_wrap.inc into g++ on stdin.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
typedef unsigned int UINT;
typedef unsigned char BYTE;
#define ASSERT(x) ((void)0)
