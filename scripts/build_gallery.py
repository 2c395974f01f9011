"""Example gallery without Sphinx: one markdown page per script of ``blades_b200/examples`` (title and narrative from
the module docstring, the code as a block, a download link) + ``docs/examples/index.md``.  The sphinx-gallery
configuration in ``docs/source/conf.py`` builds the same gallery when Sphinx is installed (reference:
docs/source/conf.py:71-76).  ``python scripts/build_gallery.py [--check]``."""
import ast
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "blades_b200", "examples")
DST = os.path.join(ROOT, "docs", "examples")


def page(name: str):
    text = open(os.path.join(SRC, name)).read()
    tree = ast.parse(text)
    doc = ast.get_docstring(tree) or name
    title = doc.strip().splitlines()[0].strip().rstrip(".")
    body = "\n".join(doc.strip().splitlines()[1:]).strip()
    code = text
    if tree.body and isinstance(tree.body[0], ast.Expr) and isinstance(getattr(tree.body[0], "value", None), ast.Constant):
        code = "\n".join(text.splitlines()[tree.body[0].end_lineno:]).lstrip("\n")
    md = f"# {title}\n\n{body}\n\n```python\n{code.rstrip()}\n```\n\n" \
         f"Source: [`blades_b200/examples/{name}`](../../blades_b200/examples/{name}) -- run it with " \
         f"`python -m blades_b200.examples.{name[:-3]}`.\n"
    return title, md


def main() -> int:
    os.makedirs(DST, exist_ok=True)
    entries = []
    for name in sorted(os.listdir(SRC)):
        if not name.endswith(".py") or name.startswith("_"):
            continue
        title, md = page(name)
        out = os.path.join(DST, name[:-3] + ".md")
        if "--check" in sys.argv:
            if not os.path.exists(out) or open(out).read() != md:
                print("stale:", out)
                return 1
        else:
            open(out, "w").write(md)
        entries.append((name[:-3], title))
    index = "# Example gallery\n\nRunnable scripts of `blades_b200/examples` (the reference's gallery, " \
            "`src/blades/examples`, with the unfinished `todo_*` scripts completed).\n\n" + \
            "\n".join(f"* [{t}]({n}.md)" for n, t in entries) + "\n"
    if "--check" in sys.argv:
        return 0 if open(os.path.join(DST, "index.md")).read() == index else 1
    open(os.path.join(DST, "index.md"), "w").write(index)
    print(f"{len(entries)} gallery pages -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
