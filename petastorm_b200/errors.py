"""Exception types shared across the package (names follow petastorm/errors.py:16 and
petastorm/etl/dataset_metadata.py:38-49 so that user ``except`` clauses keep working)."""


class NoDataAvailableError(Exception):
    """The requested sharding leaves this reader with nothing to read (petastorm/errors.py:16)."""


class PetastormMetadataError(Exception):
    """The dataset carries no (or unusable) Petastorm metadata (petastorm/etl/dataset_metadata.py:38-42)."""


class PetastormMetadataGenerationError(Exception):
    """Metadata could not be generated (petastorm/etl/dataset_metadata.py:45-49)."""


class DecodeFieldError(RuntimeError):
    """A field of a row could not be decoded (petastorm/utils.py:48)."""
