"""TEST-ONLY stand-in for openpyxl (see tests/stubs/README.md): records cells, `save` writes them as tab-separated text."""
__version__ = "3.1.5"
from .workbook import Workbook  # noqa: E402,F401


def load_workbook(*args, **kwargs):
    raise NotImplementedError("test stub: reading workbooks is not needed by dps.py")
