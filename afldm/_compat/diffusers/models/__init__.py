from . import attention_processor  # noqa: F401
