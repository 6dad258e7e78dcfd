"""Analysis tools on top of the supervisor (reference: guardians/)."""
