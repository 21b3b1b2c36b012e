from teochat_b200.eval.inference import (extract_bboxes, replace_video_token, run_inference,  # noqa: F401
                                         run_inference_batch, run_inference_single)
