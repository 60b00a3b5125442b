"""Zero-edit drop-in for the reference's drivers.

The reference's CLIs import the pipeline as ``from genpercept import GenPerceptPipeline``
(/root/reference/run.py:33, /root/reference/infer.py) and ``genpercept/__init__.py:18`` resolves that through
``from .genpercept_pipeline import GenPerceptPipeline, GenPerceptOutput``.  Python looks a submodule up in
``sys.modules`` before it looks on disk, so seeding ``sys.modules["genpercept.genpercept_pipeline"]`` with a module
that exports this package's classes makes the unmodified ``run.py`` / ``infer.py`` construct the native engine — while
``genpercept.models.*`` / ``genpercept.util.*`` (which run.py also imports) keep coming from the reference checkout.

    cd /path/to/GenPercept
    PYTHONPATH=/path/to/genpercept_b200_repo python -m genpercept_b200.dropin run.py --checkpoint ... --unet ...

or, from Python, ``import genpercept_b200.dropin; genpercept_b200.dropin.install()`` before the first
``import genpercept``.
"""
import runpy
import sys
import types

_NAME = "genpercept.genpercept_pipeline"


def install():
    """Idempotent.  Returns the seeded module."""
    mod = sys.modules.get(_NAME)
    if mod is not None and getattr(mod, "__genpercept_b200__", False):
        return mod
    from .pipeline import GenPerceptOutput, GenPerceptPipeline
    mod = types.ModuleType(_NAME, "genpercept_b200 drop-in for genpercept/genpercept_pipeline.py")
    mod.GenPerceptPipeline = GenPerceptPipeline
    mod.GenPerceptOutput = GenPerceptOutput
    mod.__genpercept_b200__ = True
    sys.modules[_NAME] = mod
    parent = sys.modules.get("genpercept")
    if parent is not None:            # the reference package was imported first: rebind what its __init__ re-exported
        parent.genpercept_pipeline = mod
        parent.GenPerceptPipeline = GenPerceptPipeline
        parent.GenPerceptOutput = GenPerceptOutput
    return mod


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        sys.exit("usage: python -m genpercept_b200.dropin <run.py|infer.py> [its arguments ...]")
    install()
    script = argv[0]
    sys.argv = argv
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))      # what `python script.py` would have done
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
