from dexb200.audio.audio_processing import dynamic_range_compression, dynamic_range_decompression  # noqa: F401
