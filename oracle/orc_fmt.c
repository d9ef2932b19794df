/*
 * oracle/orc_fmt.c -- TEST INFRASTRUCTURE ONLY (see orc_types.h).
 * Restates the columns of s_FmtConvMapping the shader chain depends on
 * (Source/Helper.cpp:295-359: plane view formats, div_chroma_w/h, Packsize,
 * PitchCoeff, CSType, Subsampling, CDepth) and the <<6 repack of
 * CopyPlane10to16 (Source/Helper.cpp:377-392, :789-803).
 */
#include <stddef.h>
#include <stdint.h>
#include "orc_api.h"

#define L_NONE 0
#define L_BI   1
#define L_P3   2
#define L_422  3
#define L_444  4
#define L_RGB  5   /* one RGBA texel per pixel; packed_kind = texture format: 0 B8G8R8X8/A8, 1 R10G10B10A2, 2 R16G16B16A16 */
#define L_GRAY 6   /* one R8 / R16 plane */

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
 { ORC_CF_V210,      L_422, 2,10, 0, 422, 2,1, 0, 2,2, 0,0, "v210"      },  /* texture = Y210 layout after orc_repack (CopyFrameV210) */
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
 /* Source/Helper.cpp:341-358.  Planes of GBRP are G, B, R (the matrix is swizzled instead, DX11VideoProcessor.cpp:863-867) */
 { ORC_CF_GBRP8,     L_P3,  1, 8, 1, 444, 1,1, 0, 1,6, 0,0, "GBRP8"     },
 { ORC_CF_GBRP10,    L_P3,  2,10, 1, 444, 1,1, 6, 2,6, 0,0, "GBRP10"    },
 { ORC_CF_GBRP16,    L_P3,  2,16, 1, 444, 1,1, 0, 2,6, 0,0, "GBRP16"    },
 /* bytes/packsize of the RGB rows describe the SAMPLE; the texture the shaders read is what orc_repack produces */
 { ORC_CF_RGB24,     L_RGB, 1, 8, 1, 444, 1,1, 0, 3,2, 0,0, "RGB24"     },
 { ORC_CF_XRGB32,    L_RGB, 1, 8, 1, 444, 1,1, 0, 4,2, 0,0, "RGB32"     },
 { ORC_CF_ARGB32,    L_RGB, 1, 8, 1, 444, 1,1, 0, 4,2, 0,0, "ARGB32"    },
 { ORC_CF_r210,      L_RGB, 4,10, 1, 444, 1,1, 0, 4,2, 0,1, "r210"      },
 { ORC_CF_RGB48,     L_RGB, 2,16, 1, 444, 1,1, 0, 6,2, 0,2, "RGB48"     },
 { ORC_CF_BGR48,     L_RGB, 2,16, 1, 444, 1,1, 0, 6,2, 0,2, "BGR48"     },
 { ORC_CF_BGRA64,    L_RGB, 2,16, 1, 444, 1,1, 0, 8,2, 0,2, "BGRA64"    },
 { ORC_CF_B64A,      L_RGB, 2,16, 1, 444, 1,1, 0, 8,2, 0,2, "b64a"      },
 { ORC_CF_Y8,        L_GRAY,1, 8, 2, 400, 1,1, 0, 1,2, 0,0, "Y8"        },
 { ORC_CF_Y10,       L_GRAY,2,10, 2, 400, 1,1, 6, 2,2, 0,0, "Y10"       },
 { ORC_CF_Y16,       L_GRAY,2,16, 2, 400, 1,1, 0, 2,2, 0,0, "Y16"       },
};

const orc_fmt* orc_get_fmt(int cformat)
{
    if (cformat < 0 || cformat >= (int)(sizeof(s_fmts)/sizeof(s_fmts[0]))) return NULL;
    if (s_fmts[cformat].layout == L_NONE) return NULL;
    return &s_fmts[cformat];
}

/* ---- sample -> source texture repacks, restating the CopyFrame* family the reference installs as
 *      m_pCopyPlaneFn (GetCopyPlaneFunction, Source/Helper.cpp:377-409).  dst is the mapped texture:
 *      RGB24 -> B8G8R8X8 (:441-470), RGB48/BGR48/BGRA64/b64a -> R16G16B16A16 (:528-665),
 *      r210 -> R10G10B10A2 (:769-787), v210 -> Y210-layout R16G16B16A16 macro-pixels (:697-738).
 *      A negative src_pitch walks a bottom-up DIB from its last row (MemCopyToTexSrcVideo, DX11VideoProcessor.cpp:1246).
 *      Returns 0, or -1 when the format has no repack (copy as is). */
static uint16_t rd_le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static void wr_le16(uint8_t* p, uint16_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

int orc_repack(int cformat, int width, int lines, const uint8_t* src, int src_pitch, uint8_t* dst, int dst_pitch)
{
    if (src_pitch < 0) src = src + (ptrdiff_t)src_pitch * (1 - lines);
    for (int y = 0; y < lines; y++) {
        const uint8_t* s = src + (ptrdiff_t)y * src_pitch;
        uint8_t* d = dst + (size_t)y * dst_pitch;
        switch (cformat) {
        case ORC_CF_RGB24:                       /* B G R bytes -> B G R X */
            for (int x = 0; x < width; x++) { d[4*x] = s[3*x]; d[4*x+1] = s[3*x+1]; d[4*x+2] = s[3*x+2]; d[4*x+3] = 0xFF; }
            break;
        case ORC_CF_XRGB32: case ORC_CF_ARGB32:  /* CopyPlaneAsIs */
            for (int x = 0; x < width * 4; x++) d[x] = s[x];
            break;
        case ORC_CF_RGB48:                       /* R G B words -> R G B A(=don't care) */
            for (int x = 0; x < width; x++) { for (int k = 0; k < 3; k++) wr_le16(d + 8*x + 2*k, rd_le16(s + 6*x + 2*k)); wr_le16(d + 8*x + 6, 0xFFFF); }
            break;
        case ORC_CF_BGR48:                       /* B G R words -> R G B */
            for (int x = 0; x < width; x++) { for (int k = 0; k < 3; k++) wr_le16(d + 8*x + 2*k, rd_le16(s + 6*x + 2*(2-k))); wr_le16(d + 8*x + 6, 0xFFFF); }
            break;
        case ORC_CF_BGRA64:                      /* swap words 0 and 2 */
            for (int x = 0; x < width; x++) { wr_le16(d + 8*x, rd_le16(s + 8*x + 4)); wr_le16(d + 8*x + 2, rd_le16(s + 8*x + 2));
                                              wr_le16(d + 8*x + 4, rd_le16(s + 8*x)); wr_le16(d + 8*x + 6, rd_le16(s + 8*x + 6)); }
            break;
        case ORC_CF_B64A:                        /* big-endian A R G B -> little-endian R G B A */
            for (int x = 0; x < width; x++) { const uint8_t* p = s + 8*x;
                wr_le16(d + 8*x,     (uint16_t)((p[2] << 8) | p[3])); wr_le16(d + 8*x + 2, (uint16_t)((p[4] << 8) | p[5]));
                wr_le16(d + 8*x + 4, (uint16_t)((p[6] << 8) | p[7])); wr_le16(d + 8*x + 6, (uint16_t)((p[0] << 8) | p[1])); }
            break;
        case ORC_CF_r210:                        /* big-endian x:2 R:10 G:10 B:10 -> R10G10B10A2 with A = 0 */
            for (int x = 0; x < width; x++) { const uint8_t* p = s + 4*x;
                const uint32_t be = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
                const uint32_t v = ((be >> 20) & 1023u) | (((be >> 10) & 1023u) << 10) | ((be & 1023u) << 20);
                d[4*x] = (uint8_t)v; d[4*x+1] = (uint8_t)(v >> 8); d[4*x+2] = (uint8_t)(v >> 16); d[4*x+3] = (uint8_t)(v >> 24); }
            break;
        case ORC_CF_V210: {                      /* 4 words = 6 pixels: Cb0 Y0 Cr0 | Y1 Cb1 Y2 | Cr1 Y3 Cb2 | Y4 Cr2 Y5 */
            for (int x = 0; x < width; x += 2) {                     /* one Y210 macro-pixel (Y0 U Y1 V), codes << 6 */
                const int g = x / 6, j = (x % 6) / 2;
                uint32_t w[4];
                for (int k = 0; k < 4; k++) { const uint8_t* p = s + 16*g + 4*k; w[k] = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
                uint32_t y0, y1, cb, cr;
                if (j == 0)      { cb = w[0];       y0 = w[0] >> 10; cr = w[0] >> 20; y1 = w[1]; }
                else if (j == 1) { y0 = w[1] >> 20; cb = w[1] >> 10; y1 = w[2] >> 10; cr = w[2]; }
                else             { cb = w[2] >> 20; y0 = w[3];       cr = w[3] >> 10; y1 = w[3] >> 20; }
                uint8_t* q = d + 4 * x;          /* 8 bytes per macro-pixel = 4 bytes per pixel */
                wr_le16(q, (uint16_t)((y0 & 1023u) << 6)); wr_le16(q + 2, (uint16_t)((cb & 1023u) << 6));
                wr_le16(q + 4, (uint16_t)((y1 & 1023u) << 6)); wr_le16(q + 6, (uint16_t)((cr & 1023u) << 6));
            }
            break; }
        default: return -1;
        }
    }
    return 0;
}
