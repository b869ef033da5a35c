"""Run the reference's train scripts against this repository WITHOUT editing them.

    python -m gshell_b200.dropin /path/to/GShell/train_gshelltet_deepfashion.py --config configs/deepfashion_mc_256.json ...

The scripts import `geometry.*`, `render.*`, `denoiser.*` as top-level packages (train_gshelltet_deepfashion.py:22-39).  `install()`
binds those names to the packages of this repository -- every module once, under both names, so that `render.mesh.Mesh` and
`gshell_b200.render.mesh.Mesh` are the same class -- and appends the reference's own `render/` directory to the search path of the
`render` package: modules that exist here win; the ones that are not part of the hot path and are not rebuilt (`render.obj`,
`render.material`, `render.texture`: OBJ / MTL / texture IO) resolve to the reference's files and see this repository's `mesh`,
`util`, `mlptexture` through their relative imports.  Everything else the scripts import (`dataset.*`, nvdiffrast for the texture
IO, xatlas) comes from the reference checkout / the environment as before.  tests/test_dropin_imports_cpu.py executes the import
block of the train scripts this way."""
import importlib
import os
import pkgutil
import runpy
import sys

PACKAGES = ("geometry", "render", "denoiser")


def install(reference_root=None):
    """Alias the hot-path packages under the reference's top-level names; with `reference_root`, also make the reference's
    remaining `render` modules and its other top-level packages importable.  Idempotent."""
    for pkg in PACKAGES:
        mod = importlib.import_module(f"gshell_b200.{pkg}")
        own = list(mod.__path__[:1])
        sys.modules[pkg] = mod
        for info in pkgutil.walk_packages(own, prefix=f"gshell_b200.{pkg}."):
            sys.modules[info.name[len("gshell_b200."):]] = importlib.import_module(info.name)
    if reference_root is not None:
        reference_root = os.path.abspath(reference_root)
        ref_render = os.path.join(reference_root, "render")
        render = sys.modules["render"]
        if os.path.isdir(ref_render) and ref_render not in render.__path__:
            render.__path__.append(ref_render)
        if reference_root not in sys.path:
            sys.path.append(reference_root)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        raise SystemExit("usage: python -m gshell_b200.dropin <reference train script> [its arguments]")
    script = os.path.abspath(argv[0])
    install(os.path.dirname(script))
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
