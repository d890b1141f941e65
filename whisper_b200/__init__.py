"""whisper_b200 - B200-native Whisper inference hot path behind the openai/whisper Python surface."""
__version__ = "0.1.0"
