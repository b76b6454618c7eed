"""Copy the reference's Python package to baseline/_ref/ (see baseline/README.md).  Build container only."""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('MEGA_NERF_REFERENCE', '/root/reference')
DST = os.path.join(HERE, '_ref')


def make(force: bool = False) -> str | None:
    src = os.path.join(SRC, 'mega_nerf')
    if not os.path.isdir(src):
        return DST if os.path.isdir(os.path.join(DST, 'mega_nerf')) else None
    dst = os.path.join(DST, 'mega_nerf')
    if os.path.isdir(dst) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    # the two scripts of the hot path's callers that tests drive (cluster masks, container merge) live outside the package
    sdir = os.path.join(SRC, 'scripts')
    if os.path.isdir(sdir):
        shutil.copytree(sdir, os.path.join(DST, 'scripts'), ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    return DST


if __name__ == '__main__':
    print(make(force='--force' in sys.argv))
