from teochat_b200.mm_utils import KeywordsStoppingCriteria, get_model_name_from_path, tokenizer_image_token  # noqa: F401
