"""Python mirror of phaneron's src/process operator surface, bound to the CUDA library."""
from .packer import Interlace, PackImpl, Packer  # noqa: F401
from .io import ToRGBA, FromRGBA  # noqa: F401
from .image_process import ImageProcess, ProcessImpl  # noqa: F401
from .combine import Combine  # noqa: F401
from .transition import Transition  # noqa: F401
from .transform import Transform  # noqa: F401
from .yadif import Yadif  # noqa: F401
from .mix import Mix  # noqa: F401
from .wipe import Wipe  # noqa: F401
from .resize import Resize  # noqa: F401
