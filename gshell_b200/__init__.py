"""gshell_b200 -- B200-native (sm_100a) implementation of the G-Shell inverse-rendering hot path.

Sub-packages mirror the reference's layout; `gshell_b200.dropin` binds them to the reference's top-level
package names so that its train scripts run unmodified:
    geometry/   GShell_Tets, GShellFlexiCubes, *Geometry.getMesh()
    render/     render.py, renderutils, optixutils, light, mesh
    denoiser/   BilateralDenoiser
All hot operators call hand-written CUDA through the C ABI in include/gshell_b200.h.
"""
__version__ = "0.1.0"
