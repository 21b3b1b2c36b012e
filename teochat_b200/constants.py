"""Model constants of the hot path (mirror of videollava/constants.py:9-24)."""

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200          # constants.py:9
DEFAULT_IMAGE_TOKEN = "<image>"   # constants.py:10
DEFAULT_VIDEO_TOKEN = "<video>"   # constants.py:17
MAX_IMAGE_LENGTH = 16             # constants.py:24

# OpenAI CLIP statistics used by the image processor (processing_image.py:7-8)
OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)
