"""TEST INFRASTRUCTURE.  Compiles ONE function / method of the reference from its own source text.

Most of the reference's hot-path modules cannot be imported in this container (mmcv, mmdet and
spconv-2.x are absent), but several functions on the path are plain torch / numpy / numba code.
`load_def` reads such a function's `def` block from the file where it lies under /root/reference at
run time, dedents it and executes it in a namespace the caller supplies -- the reference's code runs
unmodified and nothing of it is copied into this repository.
"""
import os
import textwrap

REFERENCE_ROOT = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'mmdet3d'))


def load_def(relpath, name, namespace, prefix='', keyword='def'):
    """-> the function object `name` defined in REFERENCE_ROOT/relpath (top level or method); with
    keyword='class', the class `name` (its own decorators are left out, those of its methods are not).
    `prefix` is source text put in front of the block (imports, a decorator line)."""
    lines = open(os.path.join(REFERENCE_ROOT, relpath)).read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.lstrip().startswith('%s %s(' % (keyword, name)))
    indent = len(lines[start]) - len(lines[start].lstrip())
    end = len(lines)
    for i in range(start + 1, len(lines)):
        ln = lines[i]
        if ln.strip() and len(ln) - len(ln.lstrip()) <= indent and not ln.lstrip().startswith('#'):
            end = i
            break
    ns = dict(namespace)
    exec(prefix + textwrap.dedent('\n'.join(lines[start:end])), ns)
    return ns[name]
