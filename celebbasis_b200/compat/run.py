"""python -m celebbasis_b200.compat.run <script.py> [args...]: run a reference driver script with the stand-ins for the
missing third-party packages registered and this repository's `ldm` mirror ahead of the reference's on sys.path."""
import os
import runpy
import sys


def overlay_reference_tree(ref_ldm_dir):
    """Host-side parts of the reference's own `ldm` tree that this package does not mirror (ldm.data.*: datasets and
    augmentations) stay importable: every mirrored package's search path is extended with the matching
    directory of the reference checkout the script lives in.  Mirrored modules win (they come first on __path__)."""
    import importlib
    import pkgutil
    if not os.path.isdir(ref_ldm_dir):
        return
    import ldm
    pkgs = [ldm] + [importlib.import_module(m.name) for m in pkgutil.walk_packages(ldm.__path__, "ldm.") if m.ispkg]
    for pkg in pkgs:
        rel = pkg.__name__.split(".")[1:]
        cand = os.path.join(ref_ldm_dir, *rel)
        if os.path.isdir(cand) and cand not in list(pkg.__path__):
            pkg.__path__.append(cand)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m celebbasis_b200.compat.run <script.py> [args...]")
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if root not in sys.path:
        sys.path.insert(0, root)
    from celebbasis_b200 import compat
    compat.install(verbose=True)
    import ldm  # noqa: F401  (the mirror: claims the `ldm` name before the script's own directory can)
    script = argv[0]
    sys.argv = argv
    sdir = os.path.dirname(os.path.abspath(script))
    for cand in (sdir, os.path.dirname(sdir)):
        overlay_reference_tree(os.path.join(cand, "ldm"))
    if sdir not in sys.path:
        sys.path.append(sdir)               # the script's siblings (e.g. evaluation/, main.py) stay importable, after ours
    parent = os.path.dirname(sdir)
    if parent not in sys.path:
        sys.path.append(parent)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
