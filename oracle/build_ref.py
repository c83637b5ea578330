"""Recipe for oracle/_ref: the reference's OWN similarity.py / util.py, made executable here.

TEST INFRASTRUCTURE ONLY (see oracle/similarity_oracle.py's header for who may import oracle/).

The reference is Python-2-only source and there is no python2 in this image, so its text cannot be
imported as it lies.  This script reads the sources WHERE THEY LIE under /root/reference, applies
three mechanical, semantics-preserving rewrites in memory, compiles the result and stores only the
BYTECODE (marshal) under oracle/_ref/ -- the analogue of compiling a C reference into a .so.
No reference source text is written anywhere in the repository; oracle/_ref/ is git-ignored and
travels to the GPU box with the snapshot (it is not gpurun-ignored).

The three rewrites (nothing else is touched; every function body stays the reference's):
  1. str.expandtabs(8)              -- Python 2 treats a tab as "up to the next multiple of 8";
                                       similarity.py mixes tabs and spaces (e.g. :27, :71)
  2. print statement -> print(...)  -- `print "x"` -> `print("x")`, a trailing comma becomes end=' '
  3. `from sets import Set`         -- -> `Set = set` (sets.Set is the pre-2.4 spelling of set)

The SNAP binding the reference imports (`snap`, a stripped binary: .MISSING_LARGE_BLOBS) is NOT part
of this build: oracle/ref_runner.py supplies a stand-in for its four call sites at load time.

Run:  python oracle/build_ref.py          (needs /root/reference; a no-op message otherwise)
"""
import hashlib
import json
import marshal
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get('BLP_REFERENCE_ROOT', '/root/reference')
OUT_DIR = os.path.join(HERE, '_ref')
FILES = ('similarity.py', 'util.py')


def _balanced(text):
    """True when every bracket opened in `text` is closed (string literals skipped)."""
    depth, i, n = 0, 0, len(text)
    while i < n:
        c = text[i]
        if c in '\'"':
            q = c
            i += 1
            while i < n and text[i] != q:
                i += 2 if text[i] == '\\' else 1
        elif c in '([{':
            depth += 1
        elif c in ')]}':
            depth -= 1
        elif c == '#':
            break
        i += 1
    return depth <= 0


_PRINT = re.compile(r'^(\s*)print\s+(?!\()(.*)$')


def print_statements_to_calls(src):
    """Rewrite 2: every `print <expr-list>` statement becomes a print(...) call."""
    lines, out, i = src.split('\n'), [], 0
    while i < len(lines):
        m = _PRINT.match(lines[i])
        if not m:
            out.append(lines[i])
            i += 1
            continue
        indent, body = m.group(1), m.group(2)
        while not _balanced(body) and i + 1 < len(lines):      # statement continues on the next line
            i += 1
            body += '\n' + lines[i]
        body = body.rstrip()
        if body.endswith(','):                                  # Python 2: suppress the newline
            body = body[:-1].rstrip() + ", end=' '"
        out.append('%sprint(%s)' % (indent, body))
        i += 1
    return '\n'.join(out)


def to_python3(src):
    src = src.expandtabs(8)                                                    # rewrite 1
    src = print_statements_to_calls(src)                                       # rewrite 2
    src = re.sub(r'^from sets import Set\s*$', 'Set = set', src, flags=re.M)   # rewrite 3
    return src


def build(ref_root=REF_ROOT, out_dir=OUT_DIR, quiet=False):
    """Compile the reference's path files to bytecode under oracle/_ref/.  Returns True if built."""
    if not os.path.isdir(ref_root):
        if not quiet:
            print('oracle/_ref: %s is not present here; keeping whatever is already built' % ref_root)
        return False
    os.makedirs(out_dir, exist_ok=True)
    manifest = {'python': '%d.%d' % sys.version_info[:2], 'reference_root': ref_root,
                'rewrites': ['expandtabs(8)', 'print statement -> print()', 'from sets import Set -> Set = set'],
                'files': {}}
    for name in FILES:
        with open(os.path.join(ref_root, name), 'rb') as fh:
            raw = fh.read()
        code = compile(to_python3(raw.decode('utf-8')), '<reference:%s>' % name, 'exec')
        with open(os.path.join(out_dir, name[:-3] + '.marshal'), 'wb') as fh:
            marshal.dump(code, fh)
        manifest['files'][name] = {'sha256': hashlib.sha256(raw).hexdigest(), 'bytes': len(raw)}
    with open(os.path.join(out_dir, 'MANIFEST.json'), 'w') as fh:
        json.dump(manifest, fh, indent=1)
    if not quiet:
        print('oracle/_ref: built %s from %s' % (', '.join(FILES), ref_root))
    return True


if __name__ == '__main__':
    build()
