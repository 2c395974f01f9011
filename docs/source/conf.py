"""Sphinx configuration of the blades_b200 documentation (counterpart of the reference's docs/source/conf.py:41-80).

The pages themselves are the markdown files one level up (``docs/*.md``) plus the example gallery; this file only wires
them into Sphinx.  Every optional extension is probed first, so ``sphinx-build docs/source docs/_build/html`` works with
a bare Sphinx install and gets richer (gallery, copy buttons, pydata theme) when the extras of ``docs/requirements.txt``
are present.  Without Sphinx at all, ``scripts/update_doc.sh`` renders the same pages and ``scripts/build_gallery.py``
the same gallery with the standard library + markdown_it."""
import importlib.util
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

project = "blades_b200"
author = "blades_b200 developers"
copyright = "2026, " + author


def _have(mod: str) -> bool:
    return importlib.util.find_spec(mod) is not None


extensions = ["sphinx.ext.autodoc", "sphinx.ext.autosummary", "sphinx.ext.napoleon", "sphinx.ext.viewcode",
              "sphinx.ext.mathjax", "sphinx.ext.intersphinx", "sphinx.ext.doctest"]
for _ext, _mod in (("myst_parser", "myst_parser"), ("m2r2", "m2r2"), ("sphinx_copybutton", "sphinx_copybutton"),
                   ("sphinx_gallery.gen_gallery", "sphinx_gallery")):
    if _have(_mod) and not (_ext == "m2r2" and "myst_parser" in extensions):
        extensions.append(_ext)

source_suffix = {".rst": "restructuredtext", ".md": "markdown"} if ("myst_parser" in extensions or "m2r2" in extensions) \
    else {".rst": "restructuredtext"}
napoleon_use_param = True
autodoc_mock_imports = ["ray"]                     # the reference's runtime; blades_b200 does not need it
autodoc_default_options = {"members": True, "undoc-members": False, "show-inheritance": True}
intersphinx_mapping = {"python": ("https://docs.python.org/3/", None), "numpy": ("https://numpy.org/doc/stable/", None),
                       "torch": ("https://pytorch.org/docs/stable", None)}
templates_path = ["_templates"]
exclude_patterns = ["_build"]

# example gallery: the runnable scripts of blades_b200/examples (the reference skips its unfinished ``todo_*`` ones,
# conf.py:71-76; here they are finished, so nothing is skipped)
sphinx_gallery_conf = {
    "line_numbers": False,
    "examples_dirs": os.path.join(ROOT, "blades_b200", "examples"),
    "gallery_dirs": "examples",
    "filename_pattern": r"/plot_",                 # only plot_* scripts are executed at build time
    "plot_gallery": "False" if os.environ.get("BLADES_DOCS_NO_RUN") else "True",
}

html_theme = "pydata_sphinx_theme" if _have("pydata_sphinx_theme") else "alabaster"
html_static_path = []
html_theme_options = {"navigation_depth": 5, "collapse_navigation": False} if html_theme == "pydata_sphinx_theme" else {}
