/* stand-in: errcodes not needed */
