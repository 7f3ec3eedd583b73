// The pixel-format conversions the reference's texture loaders lean on (product host code).
//
// scene/src/texture/loader.rs:43-87 keeps Rgb8 / Rgb32F (resp. Luma8 / LumaA8) images as they are and sends everything else through
// DynamicImage::to_rgb8() resp. to_luma8(); EnvironmentLight::new (primitive/impls/environment_light.rs:36-37) calls to_rgb32f().  Those
// are the `image` crate's (Cargo.lock: image 0.25.6, not vendored in the reference tree) generic colour conversions, restated here
// from its published rules for an already DECODED buffer (the decoder itself -- PNG, OpenEXR -- is the host's):
//   sample depth   u16 -> u8: (c + 128) / 257        u8 / u16 -> f32: c / MAX        f32 -> u8 / u16: round(clamp(c, 0, 1) * MAX)
//   RGB -> luma    (2126 r + 7152 g + 722 b) / 10000 in the source sample type: truncating u32 arithmetic for u8 / u16, f64 for f32;
//                  the depth conversion comes AFTER it
//   luma -> RGB    the (depth-converted) value three times;  alpha is dropped
#pragma once
#include <cmath>
#include <cstdint>

namespace tcpt {

enum { IMG_U8 = 0, IMG_U16 = 1, IMG_F32 = 2 };
enum { IMG_TO_RGB8 = 0, IMG_TO_LUMA8 = 1, IMG_TO_RGB32F = 2 };

namespace image_detail {
inline uint8_t u16_to_u8(uint16_t c) { return (uint8_t)(((uint32_t)c + 128u) / 257u); }
inline uint8_t f32_to_u8(float c) {
    const float v = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);   // f32::clamp; NaN passes through and NumCast of NaN fails in the crate: mapped to 0 here
    const float r = std::round(v * 255.0f);
    return r != r ? (uint8_t)0 : (uint8_t)r;
}
template <typename T> struct Sample;
template <> struct Sample<uint8_t> {
    static uint8_t luma(const uint8_t* p) { return (uint8_t)((2126u * p[0] + 7152u * p[1] + 722u * p[2]) / 10000u); }
    static uint8_t to_u8(uint8_t c) { return c; }
    static float to_f32(uint8_t c) { return (float)c / 255.0f; }
};
template <> struct Sample<uint16_t> {
    static uint16_t luma(const uint16_t* p) { return (uint16_t)((2126u * p[0] + 7152u * p[1] + 722u * p[2]) / 10000u); }
    static uint8_t to_u8(uint16_t c) { return u16_to_u8(c); }
    static float to_f32(uint16_t c) { return (float)c / 65535.0f; }
};
template <> struct Sample<float> {
    static float luma(const float* p) { return (float)((2126.0 * (double)p[0] + 7152.0 * (double)p[1] + 722.0 * (double)p[2]) / 10000.0); }
    static uint8_t to_u8(float c) { return f32_to_u8(c); }
    static float to_f32(float c) { return c; }
};
template <typename T>
void convert(const T* src, size_t n_pix, uint32_t ch, int dst_kind, void* dst) {
    const bool colour = ch >= 3;
    for (size_t i = 0; i < n_pix; ++i) {
        const T* p = src + i * ch;
        if (dst_kind == IMG_TO_LUMA8) {
            ((uint8_t*)dst)[i] = Sample<T>::to_u8(colour ? Sample<T>::luma(p) : p[0]);
        } else if (dst_kind == IMG_TO_RGB8) {
            uint8_t* o = (uint8_t*)dst + 3 * i;
            if (colour) { o[0] = Sample<T>::to_u8(p[0]); o[1] = Sample<T>::to_u8(p[1]); o[2] = Sample<T>::to_u8(p[2]); }
            else o[0] = o[1] = o[2] = Sample<T>::to_u8(p[0]);
        } else {
            float* o = (float*)dst + 3 * i;
            if (colour) { o[0] = Sample<T>::to_f32(p[0]); o[1] = Sample<T>::to_f32(p[1]); o[2] = Sample<T>::to_f32(p[2]); }
            else o[0] = o[1] = o[2] = Sample<T>::to_f32(p[0]);
        }
    }
}
}  // namespace image_detail

// src: height * width * channels samples (channels 1 = L, 2 = LA, 3 = RGB, 4 = RGBA).  false on a bad argument.
inline bool image_convert(const void* src, uint32_t width, uint32_t height, uint32_t channels, int sample_type, int dst_kind, void* dst) {
    if (!src || !dst || channels < 1 || channels > 4 || dst_kind < 0 || dst_kind > 2) return false;
    const size_t n = (size_t)width * height;
    if (sample_type == IMG_U8) image_detail::convert((const uint8_t*)src, n, channels, dst_kind, dst);
    else if (sample_type == IMG_U16) image_detail::convert((const uint16_t*)src, n, channels, dst_kind, dst);
    else if (sample_type == IMG_F32) image_detail::convert((const float*)src, n, channels, dst_kind, dst);
    else return false;
    return true;
}

}  // namespace tcpt
