from . import transducer  # noqa: F401
from . import decoding_cpp  # noqa: F401
from . import envelope  # noqa: F401
from . import decoding_cy  # noqa: F401
from . import prefix_search  # noqa: F401
from . import decode  # noqa: F401
