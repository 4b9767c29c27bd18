"""The error contract of the path: JWave's checked exception hierarchy.

Mirrors jwave/exceptions/JWaveException.java:32-101, JWaveFailure.java:32-52 and
JWaveError.java:32-52 (paths relative to the reference's src/main/java)."""
import sys


class JWaveException(Exception):
    """jwave/exceptions/JWaveException.java:32"""

    def __init__(self, message="JWave: Exception"):
        super().__init__(message)
        self._message = message

    def getMessage(self):
        return self._message

    def showMessage(self):
        """JWaveException.java:91 - prints the message"""
        print(self._message, file=sys.stdout)


class JWaveFailure(JWaveException):
    """Recoverable misuse: wrong length, level out of range (JWaveFailure.java:32)."""


class JWaveError(JWaveException):
    """Unrecoverable: here, a CUDA / NCCL failure reported by libjwave_cuda.so (JWaveError.java:32)."""
