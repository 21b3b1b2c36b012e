from teochat_b200.mm_utils import (KeywordsStoppingCriteria, expand2square, get_model_name_from_path,  # noqa: F401
                                     load_image_from_base64, process_images, tokenizer_image_token)
