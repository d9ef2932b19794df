/*
 * oracle/orc_fmt.c -- TEST INFRASTRUCTURE ONLY (see orc_types.h).
 * Restates the columns of s_FmtConvMapping the shader chain depends on
 * (Source/Helper.cpp:295-359: plane view formats, div_chroma_w/h, Packsize,
 * PitchCoeff, CSType, Subsampling, CDepth) and the <<6 repack of
 * CopyPlane10to16 (Source/Helper.cpp:377-392, :789-803).
 */
#include <stddef.h>
#include "orc_api.h"

#define L_NONE 0
#define L_BI   1
#define L_P3   2
#define L_422  3
#define L_444  4

static const orc_fmt s_fmts[] = {
 /* cformat          layout bytes cdepth cs sub  dw dh shift pack pc swap kind name */
 { ORC_CF_NONE,      L_NONE,0, 0, 0,   0, 1,1, 0, 0,0, 0,0, "unknown"   },
 { ORC_CF_NV12,      L_BI,  1, 8, 0, 420, 2,2, 0, 1,3, 0,0, "NV12"      },
 { ORC_CF_P010,      L_BI,  2,16, 0, 420, 2,2, 0, 2,3, 0,0, "P010"      },
 { ORC_CF_P016,      L_BI,  2,16, 0, 420, 2,2, 0, 2,3, 0,0, "P016"      },
 { ORC_CF_YUY2,      L_422, 1, 8, 0, 422, 2,1, 0, 2,2, 0,0, "YUY2"      },
 { ORC_CF_UYVY,      L_422, 1, 8, 0, 422, 2,1, 0, 2,2, 0,1, "UYVY"      },
 { ORC_CF_P210,      L_BI,  2,16, 0, 422, 2,1, 0, 2,4, 0,0, "P210"      },
 { ORC_CF_P216,      L_BI,  2,16, 0, 422, 2,1, 0, 2,4, 0,0, "P216"      },
 { ORC_CF_Y210,      L_422, 2,10, 0, 422, 2,1, 0, 4,2, 0,0, "Y210"      },
 { ORC_CF_Y216,      L_422, 2,16, 0, 422, 2,1, 0, 4,2, 0,0, "Y216"      },
 { ORC_CF_V210,      L_NONE,0,10, 0, 422, 2,1, 0, 2,2, 0,0, "v210"      },  /* repack: next row */
 { ORC_CF_AYUV,      L_444, 1, 8, 0, 444, 1,1, 0, 4,2, 0,0, "AYUV"      },
 { ORC_CF_Y410,      L_444, 4,10, 0, 444, 1,1, 0, 4,2, 0,1, "Y410"      },
 { ORC_CF_Y416,      L_444, 2,16, 0, 444, 1,1, 0, 8,2, 0,2, "Y416"      },
 { ORC_CF_YV12,      L_P3,  1, 8, 0, 420, 2,2, 0, 1,3, 1,0, "YV12"      },
 { ORC_CF_YV16,      L_P3,  1, 8, 0, 422, 2,1, 0, 1,4, 1,0, "YV16"      },
 { ORC_CF_YV24,      L_P3,  1, 8, 0, 444, 1,1, 0, 1,6, 1,0, "YV24"      },
 { ORC_CF_YUV420P8,  L_P3,  1, 8, 0, 420, 2,2, 0, 1,3, 0,0, "YUV420P8"  },
 { ORC_CF_YUV422P8,  L_P3,  1, 8, 0, 422, 2,1, 0, 1,4, 0,0, "YUV422P8"  },
 { ORC_CF_YUV444P8,  L_P3,  1, 8, 0, 444, 1,1, 0, 1,6, 0,0, "YUV444P8"  },
 { ORC_CF_YUV420P10, L_P3,  2,10, 0, 420, 2,2, 6, 2,3, 0,0, "YUV420P10" },
 { ORC_CF_YUV420P16, L_P3,  2,16, 0, 420, 2,2, 0, 2,3, 0,0, "YUV420P16" },
 { ORC_CF_YUV422P10, L_P3,  2,10, 0, 422, 2,1, 6, 2,4, 0,0, "YUV422P10" },
 { ORC_CF_YUV422P16, L_P3,  2,16, 0, 422, 2,1, 0, 2,4, 0,0, "YUV422P16" },
 { ORC_CF_YUV444P10, L_P3,  2,10, 0, 444, 1,1, 6, 2,6, 0,0, "YUV444P10" },
 { ORC_CF_YUV444P16, L_P3,  2,16, 0, 444, 1,1, 0, 2,6, 0,0, "YUV444P16" },
};

const orc_fmt* orc_get_fmt(int cformat)
{
    if (cformat < 0 || cformat >= (int)(sizeof(s_fmts)/sizeof(s_fmts[0]))) return NULL;
    if (s_fmts[cformat].layout == L_NONE) return NULL;
    return &s_fmts[cformat];
}
