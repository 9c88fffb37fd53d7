"""TEST INFRASTRUCTURE ONLY -- byte-compile the reference's Python modules from where they lie into one zip archive of
sourceless .pyc files (a build output under oracle/_ref/, like the compiled extensions; no reference source is copied).
usage: zip_reference.py <reference root> <out.zip>"""
import os
import py_compile
import sys
import tempfile
import zipfile

ref, out = sys.argv[1], sys.argv[2]
with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(out, "w", zipfile.ZIP_STORED) as z:
    for root, _, files in os.walk(os.path.join(ref, "poreover")):
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            src = os.path.join(root, f)
            rel = os.path.relpath(src, ref)
            cfile = os.path.join(tmp, "m.pyc")
            py_compile.compile(src, cfile=cfile, dfile="reference/" + rel, doraise=True)
            z.write(cfile, rel[:-3] + ".pyc")
