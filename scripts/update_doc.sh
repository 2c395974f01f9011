#!/bin/bash
# Render docs/*.md to docs/_build/html (plain markdown -> html with the `markdown_it` package; no Sphinx needed).
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import os
from markdown_it import MarkdownIt
md = MarkdownIt("commonmark").enable("table")
os.makedirs("docs/_build/html", exist_ok=True)
for name in sorted(os.listdir("docs")):
    if name.endswith(".md"):
        html = md.render(open(os.path.join("docs", name)).read()).replace('.md"', '.html"')
        with open(os.path.join("docs/_build/html", name[:-3] + ".html"), "w") as f:
            f.write(f"<html><head><meta charset='utf-8'><title>{name[:-3]}</title></head><body>{html}</body></html>")
        print("rendered", name)
PY
