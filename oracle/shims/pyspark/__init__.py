"""Harness-side stub of pyspark (absent from this image; no JVM).  TEST INFRASTRUCTURE ONLY: lets the unmodified
reference un-pickle the Unischema stored in ``_common_metadata`` (the pickles reference ``pyspark.sql.types.*`` and
``pyspark.serializers._restore``)."""
