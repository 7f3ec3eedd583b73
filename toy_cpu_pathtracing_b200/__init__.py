"""toy-cpu-pathtracing_b200: B200-native (sm_100a) backend for the path-integration hot path of toy-cpu-pathtracing.

The compute lives in `lib/libtcpt.so` (hand-written CUDA wavefront kernels behind the C ABI of include/tcpt.h); this
package is the host-side mirror of the reference's `scene` / `renderer` construction API on top of it.
"""
from . import capi  # noqa: F401
from .capi import Context, TcptError  # noqa: F401
from .renderer import (AlbedoRenderer, NormalRenderer, BoxFilter, Camera, RandomSampler, ReinhardToneMap, RendererArgs, RendererImage, SrgbRendererMis,  # noqa: F401
                       SrgbRendererNee, SrgbRendererPt, ZSobolSampler, RENDERERS)
from .scene import (ColorSrgb, ColorSrgbLinear, ConstantSpectrum, CreatePrimitiveDesc, EmissiveMaterial, FloatParameter,  # noqa: F401
                    FloatTexture, GlassMaterial, GlassType, LambertMaterial, MetalMaterial, MetalType, NormalParameter, NormalTexture, PlasticMaterial, RgbAlbedoSpectrum, RgbTexture,
                    Scene, SceneDescription, SimpleClearcoatPbrMaterial, SimplePbrMaterial, SpectrumParameter, SpectrumType,
                    Transform, presets)
