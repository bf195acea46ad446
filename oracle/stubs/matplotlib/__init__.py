"""Import-time stand-in for matplotlib (absent in this image).

Test infrastructure only: lets `oracle/ref_runner.py` import the *reference*
package from /root/reference/src, whose modules import matplotlib at module
scope.  Nothing here draws anything.
"""
scale = None
