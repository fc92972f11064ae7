from pcseqlearning_b200.preprocessors.registration_utils import *  # noqa: F401,F403
from pcseqlearning_b200.preprocessors.registration_utils import register_to_next_frame  # noqa: F401
