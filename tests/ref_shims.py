"""The three shims SURVEY.md §7 step 1 lists for importing the reference's Runner in this image (nothing here is on the
product path): `np.int` (removed in numpy >= 1.24, used in an annotation at filesystem_dataset.py:306), the `lpips`
package (metrics.py:5; never called by render_image / _training_step) and `configargparse` (opts.py:1)."""
import argparse
import sys
import types

import numpy as np


def install_shims() -> None:
    if not hasattr(np, 'int'):
        np.int = int                                         # noqa: NPY001
    if 'lpips' not in sys.modules:
        try:
            import lpips  # noqa: F401
        except Exception:  # noqa: BLE001
            m = types.ModuleType('lpips')

            class LPIPS:                                     # placeholder: metrics.lpips() is not part of the hot path
                def __init__(self, *a, **k):
                    raise RuntimeError('lpips is not installed in this image')
            m.LPIPS = LPIPS
            sys.modules['lpips'] = m
    if 'configargparse' not in sys.modules:
        try:
            import configargparse  # noqa: F401
        except Exception:  # noqa: BLE001
            m = types.ModuleType('configargparse')

            class ArgParser(argparse.ArgumentParser):
                def __init__(self, *a, config_file_parser_class=None, **k):
                    super().__init__(*a, **k)

                def add_argument(self, *a, is_config_file=False, **k):
                    return super().add_argument(*a, **k)
            m.ArgParser = m.ArgumentParser = ArgParser
            m.YAMLConfigFileParser = object
            sys.modules['configargparse'] = m
