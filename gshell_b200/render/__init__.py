"""`render` package of gshell-b200: render.py, renderutils, optixutils, light, mesh, regularizer, mlptexture, util (the hot-path
half of the reference's `render` package; `gshell_b200.dropin` maps it onto the reference's top-level name)."""
