"""speakerguard_b200: B200-native engine for SpeakerGuard's gradient-based attack iteration."""
__version__ = "0.1.0"
