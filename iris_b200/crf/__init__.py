from .model_crf import EmorCRF, parse_emor_file      # noqa: F401
