"""TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  Never imported by selfreconcode_b200/."""
