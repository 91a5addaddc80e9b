"""Import alias: ``vqa_b200`` resolves to the sources in
``visual-question-answering-vqa-system_b200/`` (the directory name the project layout
prescribes is not a valid Python identifier, so this shim points ``__path__`` at it)."""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "visual-question-answering-vqa-system_b200")
__path__.insert(0, _SRC)

from ._api import *  # noqa: E402,F401,F403
from ._api import __all__  # noqa: E402,F401
