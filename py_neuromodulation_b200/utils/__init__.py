from . import types  # noqa: F401
from . import channels  # noqa: F401
from . import io  # noqa: F401
