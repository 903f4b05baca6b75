"""Drop-in replacement for the reference's ``model`` package (same import paths, class names,
constructor signatures, config-JSON surface and state_dict keys -- SURVEY.md section 8b / Appendix B).
Put this directory's parent (``video-captioning-transformer_b200/``) ahead of the reference tree on
``sys.path`` and the reference's train.py / eval.py / predict_video.py run unmodified on the B200
kernels of libvct_b200.so.  There is no CPU fallback: the hot path raises without CUDA."""
