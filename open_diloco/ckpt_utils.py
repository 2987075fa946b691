"""Reference ``open_diloco.ckpt_utils`` surface."""
from opendiloco_b200.utils.ckpt import (CKPT_PREFIX, GLOBAL_STATE_FILE, CkptConfig, check_checkpoint_path_access,  # noqa: F401
                                        delete_old_checkpoints, filter_ckpt_files, get_diloco_rank_dir_name, get_resume_info,
                                        load_checkpoint, save_checkpoint)
