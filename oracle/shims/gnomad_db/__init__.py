"""gnomad_db shim -- TEST INFRASTRUCTURE (oracle).  Table-driven stand-in for the sqlite-backed
package the reference imports at BaseCellCalling.step2.py:9,100-108."""
